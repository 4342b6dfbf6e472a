#!/bin/bash
# A/B of the `used` map of the one-warp-per-frame region growing: records (0) / bitmap in shared (1) / bitmap in global (2)
mkdir -p gpurun_out
export PLSLAM_GROW_MODE=0
for u in 0 1 2; do
  for v in 3 7; do
    [ $u = 0 ] && [ $v = 7 ] && continue
    echo "#### USED=$u VARIANT=$v"
    export PLSLAM_GROW_USED=$u PLSLAM_GROW_VARIANT=$v
    timeout 300 python -m pytest tests/test_lines_gpu.py -x -q 2>&1 | tail -2
    timeout 120 python tools/prof_lines.py 256
    timeout 120 python tools/prof_lines.py 16
    timeout 200 python tools/prof_lines_conc.py 16
  done
done 2>&1 | tee gpurun_out/r02_bm_ab.log
