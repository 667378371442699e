#!/bin/bash
# development aid: rebuild the resampler with -D flags on the GPU box and time a few cases per variant
for v in "$@"; do
  touch pyaudiorestoration_b200/csrc/resample.cu
  make -C pyaudiorestoration_b200/csrc EXTRA="$v" > /dev/null 2>&1 || echo "BUILD FAILED for $v"
  echo "== variant '$v': $(grep -A2 'sinc_kernel_wsILi2ELi64' pyaudiorestoration_b200/csrc/resample.ptxas.log | grep -oE 'Used [0-9]+ registers|[0-9]+ bytes spill stores' | tr '\n' ' ')"
  timeout 200 python scripts/ws_ab.py quick 2>&1 | tail -5
done
touch pyaudiorestoration_b200/csrc/resample.cu
make -C pyaudiorestoration_b200/csrc > /dev/null 2>&1
