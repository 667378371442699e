#!/bin/bash
# compute-sanitizer passes over a small slice of the GPU parity tests (the reference has
# no race detection; this is ours).  memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory
# hazards in the slot-barrier STFT, the tiled resampler and the positions tile transpose.
SEL="test_stft_golden_cases or test_stft_zeropad or test_istft_golden or test_sinc_golden or test_speed_to_pos_golden_bit_exact or test_linear_mode or test_varispeed_capacity_error or test_stft_large_zeropad_and_multichannel or test_time_shards"
for tool in memcheck racecheck; do
  compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 5 python -m pytest tests -m gpu -q -x --timeout 1500 -k "$SEL" \
      > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitizer_$tool.log | tail -3 | tr '\n' ' ')"
done
# shard entry points with exactly sized slices: without torch's caching allocator every tensor is its own cudaMalloc, so a
# read one element past a rank's positions / samples slice is an error here (round 2: the resampler's tail-period read)
PYTORCH_NO_CUDA_MEMORY_CACHING=1 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests -m gpu -q -x \
    --timeout 1500 -k "test_time_shards or test_shared_segment_sums" > gpurun_out/sanitizer_shards_memcheck.log 2>&1
echo "shards memcheck exit=$? : $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitizer_shards_memcheck.log | tail -2 | tr '\n' ' ')"
