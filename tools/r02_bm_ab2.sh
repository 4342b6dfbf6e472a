#!/bin/bash
# A/B of the `used` map and of the prefetch switches in the pipelined bench, after the instruction-cache fix of k_lsd_grow
mkdir -p gpurun_out
for cfg in "0 3" "2 3" "2 7" "0 1" "1 3"; do
  set -- $cfg
  export PLSLAM_GROW_USED=$1 PLSLAM_GROW_VARIANT=$2
  echo "#### USED=$1 VARIANT=$2"
  timeout 300 python bench.py --no-cpu-baseline --no-latency --steps 48 > gpurun_out/ab2.json 2> gpurun_out/ab2.err
  python tools/benchline.py ab2 < gpurun_out/ab2.json
done 2>&1 | tee gpurun_out/r02_bm_ab2.log
