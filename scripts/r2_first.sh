#!/bin/bash
python -m pytest tests -m gpu -q -x --timeout 900 -k "sinc or varispeed or resample or shard or chunk" > gpurun_out/pytest_sinc.log 2>&1; tail -4 gpurun_out/pytest_sinc.log
bash scripts/try_variants.sh "" "-DSINC_EXPERIMENT_SKIP_TAPS" "-DSINC_EXPERIMENT_ALL_FC1" "-DSINC_EXPERIMENT_ALL_LOWPASS" "-DSINC_MIN_BLOCKS=3" > gpurun_out/variants.log 2>&1; cat gpurun_out/variants.log
