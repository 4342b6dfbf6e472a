#!/bin/bash
# copy streams in their own priority class (no hardware-queue aliasing with the compute streams), with and without ORB_AFTER=1
mkdir -p gpurun_out
for cfg in "001 0" "001 1" "100 1"; do
  set -- $cfg
  export PLSLAM_STREAM_PRIO=$1 PLSLAM_ORB_AFTER=$2
  for st in 64 20; do
    echo "#### PRIO=$1 ORB_AFTER=$2 steps=$st"
    timeout 300 python bench.py --no-cpu-baseline --no-latency --steps $st --warmup 5 > gpurun_out/p2.json 2> gpurun_out/p2.err || tail -3 gpurun_out/p2.err
    python tools/benchline.py p2 < gpurun_out/p2.json
  done
done 2>&1 | tee gpurun_out/r02_prio2.log
