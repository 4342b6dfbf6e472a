#!/bin/bash
# stream priorities of the two branches of a slot against throughput (64 steps and the driver's 20)
mkdir -p gpurun_out
for pr in ${PRIOS:-000 100 110 101 111}; do
  export PLSLAM_STREAM_PRIO=$pr
  for st in 64 20; do
    echo "#### PRIO=$pr steps=$st"
    timeout 300 python bench.py --no-cpu-baseline --no-latency --steps $st --warmup 5 > gpurun_out/prio.json 2> gpurun_out/prio.err || tail -3 gpurun_out/prio.err
    python tools/benchline.py prio < gpurun_out/prio.json
  done
done 2>&1 | tee gpurun_out/r02_prio.log
