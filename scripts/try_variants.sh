#!/bin/bash
# development aid: rebuild the resampler with experiment macros on the GPU box and time the stages
# usage: try_variants.sh "<flags of variant 1>" "<flags of variant 2>" ...
if [ $# -eq 0 ]; then set -- "" "-DSINC_EXPERIMENT_SKIP_TAPS" "-DSINC_EXPERIMENT_ALL_FC1" "-DSINC_EXPERIMENT_ALL_LOWPASS"; fi
for v in "$@"; do
  touch pyaudiorestoration_b200/csrc/resample.cu
  make -C pyaudiorestoration_b200/csrc EXTRA="$v" > /dev/null 2>&1 || echo "BUILD FAILED for $v"
  grep -A2 "sinc_kernelILi2ELi64" pyaudiorestoration_b200/csrc/resample.ptxas.log | grep -E "registers" | head -1
  echo "variant '$v': $(python bench.py --steps 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["stage_ms"])')"
done
touch pyaudiorestoration_b200/csrc/resample.cu
make -C pyaudiorestoration_b200/csrc > /dev/null 2>&1
