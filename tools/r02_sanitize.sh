#!/bin/bash
# compute-sanitizer over every kernel of the path: memcheck (both region-growing forms, odd sizes, the matcher kernels through
# their GPU tests) and racecheck (shared-memory hazards; small frames).  Logs end in "done" / the pytest summary.
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
for mode in auto 0; do
  if [ $mode = auto ]; then unset PLSLAM_GROW_MODE; else export PLSLAM_GROW_MODE=$mode; fi
  echo "#### memcheck, PLSLAM_GROW_MODE=$mode"
  timeout 1200 $S --tool memcheck --print-limit 20 python tools/sanitize_small.py 2>&1 | tail -12
done > gpurun_out/r02_memcheck.log 2>&1
unset PLSLAM_GROW_MODE
echo "#### memcheck, matcher kernels (pytest)" >> gpurun_out/r02_memcheck.log
timeout 1500 $S --tool memcheck --print-limit 20 python -m pytest tests/test_match_gpu.py tests/test_triangulation_gpu.py tests/test_bow_gpu.py tests/test_frame_gpu.py -x -q -m gpu 2>&1 | tail -8 >> gpurun_out/r02_memcheck.log
for mode in auto 0; do
  if [ $mode = auto ]; then unset PLSLAM_GROW_MODE; else export PLSLAM_GROW_MODE=$mode; fi
  echo "#### racecheck, PLSLAM_GROW_MODE=$mode"
  SAN_SMALL=1 timeout 1500 $S --tool racecheck --print-limit 20 python tools/sanitize_small.py 2>&1 | tail -30
done > gpurun_out/r02_racecheck.log 2>&1
tail -4 gpurun_out/r02_memcheck.log; tail -6 gpurun_out/r02_racecheck.log
