#!/bin/bash
# frames per CTA of k_lsd_grow (a CTA holds its registers until its slowest frame is done)
mkdir -p gpurun_out
for gw in 2 1 4; do
  export PLSLAM_GROW_GW=$gw
  for st in 64 20; do
    echo "#### GROW_GW=$gw steps=$st"
    timeout 300 python bench.py --no-cpu-baseline --no-latency --steps $st --warmup 5 > gpurun_out/gw.json 2> gpurun_out/gw.err || tail -3 gpurun_out/gw.err
    python tools/benchline.py gw < gpurun_out/gw.json
  done
done 2>&1 | tee gpurun_out/r02_gw.log
