#!/bin/bash
# Round-end evidence run on one B200 (under gpurun): GPU tests, smoke, the bench line, the reference arm, the ncu launch
# list of the same bench command, and one ncu --set full capture of every kernel of the step.
set -x
python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -c 600 gpurun_out/bench_cfg2.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 400 gpurun_out/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --no-competitor --no-strong > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"sinc_kernel|stft_tma_kernel|expand_positions|add_offsets|segment_sums" -s 12 -c 4 -f -o gpurun_out/prof_final \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --no-competitor --no-strong > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -8
