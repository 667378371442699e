#!/bin/bash
python -m pytest tests -m gpu -q -x --timeout 1200 > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
python bench.py --steps 10 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -c 3000 gpurun_out/bench_r2b.json; tail -5 gpurun_out/bench_r2b.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2b.json 2> gpurun_out/bench_ref_r2b.err; tail -c 900 gpurun_out/bench_ref_r2b.json
