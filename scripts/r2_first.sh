#!/bin/bash
# round-2 first GPU run: parity of the rewritten sinc kernel, stage split, bench line
python -m pytest tests -m gpu -q -x --timeout 900 -k "sinc or varispeed or resample or shard or chunk or cfg" > gpurun_out/pytest_sinc.log 2>&1; tail -15 gpurun_out/pytest_sinc.log
bash scripts/try_variants.sh > gpurun_out/variants.log 2>&1; cat gpurun_out/variants.log
python bench.py --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 1500 gpurun_out/bench_r2a.json
