#!/bin/bash
# ncu --set full capture of the sinc kernel for one or two build variants: r2_prof.sh <tag> "<flags>" ...
while [ $# -ge 2 ]; do
  tag=$1; flags=$2; shift 2
  touch pyaudiorestoration_b200/csrc/resample.cu
  make -C pyaudiorestoration_b200/csrc EXTRA="$flags" > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:sinc_kernel -s 3 -c 1 -f -o gpurun_out/prof_sinc_$tag \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_sinc_$tag.log 2>&1
  tail -1 gpurun_out/ncu_sinc_$tag.log
done
touch pyaudiorestoration_b200/csrc/resample.cu
make -C pyaudiorestoration_b200/csrc > /dev/null 2>&1
