#!/bin/bash
python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
bash scripts/try_variants.sh "" "-DSINC_EXPERIMENT_SKIP_TAPS" "-DSINC_EXPERIMENT_ALL_FC1" "-DSINC_EXPERIMENT_ALL_LOWPASS" > gpurun_out/variants.log 2>&1; cat gpurun_out/variants.log
python bench.py --steps 10 --no-cpu-baseline --no-strong > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2c.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'stage', d['stage_ms'], 'e2e', d['e2e']['ms_per_step'], 'parity', d['parity']['pass'], d['competitor_torch_stft']['device_ms'])
PY
tail -3 gpurun_out/bench_r2c.err
