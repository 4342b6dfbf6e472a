#!/bin/bash
# when the ORB branch of a slot starts relative to the line branch's region-growing kernel
mkdir -p gpurun_out
for oa in 1 2; do
  export PLSLAM_ORB_AFTER=$oa
  for st in 64 20; do
    echo "#### ORB_AFTER=$oa steps=$st"
    timeout 300 python bench.py --no-cpu-baseline --no-latency --steps $st --warmup 5 > gpurun_out/oa.json 2> gpurun_out/oa.err || tail -3 gpurun_out/oa.err
    python tools/benchline.py oa < gpurun_out/oa.json
  done
done 2>&1 | tee gpurun_out/r02_orbafter.log
