#!/bin/bash
python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python scripts/roofline_sweep.py > gpurun_out/sweep_n1.md 2> gpurun_out/sweep_n1.err; cat gpurun_out/sweep_n1.md; tail -3 gpurun_out/sweep_n1.err
