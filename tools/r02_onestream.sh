#!/bin/bash
# one stream per slot (no queue aliasing at <= 30 slots) against two
mkdir -p gpurun_out
for cfg in "0 32" "1 32" "2 32" "2 28" "1 28"; do
  set -- $cfg
  export PLSLAM_ONE_STREAM=$1
  for st in 64 20; do
    echo "#### ONE_STREAM=$1 depth=$2 steps=$st"
    timeout 300 python bench.py --no-cpu-baseline --no-latency --steps $st --warmup 5 --depth $2 > gpurun_out/os.json 2> gpurun_out/os.err || tail -3 gpurun_out/os.err
    python tools/benchline.py os < gpurun_out/os.json
  done
done 2>&1 | tee gpurun_out/r02_onestream.log
