#!/bin/bash
# development aid: rebuild the resampler with different launch bounds on the GPU box and time it
for mb in 2 3 4; do
  touch pyaudiorestoration_b200/csrc/resample.cu
  make -C pyaudiorestoration_b200/csrc EXTRA=-DSINC_MIN_BLOCKS=$mb > /dev/null 2>&1
  echo "SINC_MIN_BLOCKS=$mb: $(python bench.py --steps 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["stage_ms"])')"
done
touch pyaudiorestoration_b200/csrc/resample.cu
make -C pyaudiorestoration_b200/csrc > /dev/null 2>&1
