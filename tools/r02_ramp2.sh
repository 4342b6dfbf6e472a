#!/bin/bash
# robustness of the ramped first waves (e2e leg): repeats and neighbours of "4,16" at 20 steps, and ramps at 64 steps
mkdir -p gpurun_out
export PLSLAM_ORB_AFTER=1
one() { echo "#### steps=$1 ramp=$2"; timeout 300 python bench.py --no-cpu-baseline --no-latency --steps $1 --warmup 5 --wave-ramp "$2" > gpurun_out/ramp.json 2> gpurun_out/ramp.err || tail -3 gpurun_out/ramp.err; python tools/benchline.py ramp < gpurun_out/ramp.json; }
{
one 20 "4,16"; one 20 "4,16"; one 20 "5,15"; one 20 "3,17"; one 20 "4,8,8"; one 20 "6,14"; one 20 "8,12"
one 64 "4,12"; one 64 "4,16,12"; one 64 "8,8"
} 2>&1 | tee gpurun_out/r02_ramp2.log
