#!/bin/bash
# the driver's invocation (--steps 20 --warmup 5): slots in flight and the residency of k_lsd_grow
mkdir -p gpurun_out
run() { echo "#### $*"; timeout 300 python bench.py --no-cpu-baseline --no-latency "$@" > gpurun_out/s20.json 2> gpurun_out/s20.err || tail -5 gpurun_out/s20.err; python tools/benchline.py s20 < gpurun_out/s20.json; }
{
run --steps 20 --warmup 5
run --steps 20 --warmup 5 --depth 16
run --steps 20 --warmup 5 --depth 10
export PLSLAM_LIB=$PWD/rgbd-pl-slam_b200/libplslam_b200_v9.so
echo "== 9 CTAs per SM (56 registers, 832-entry shared list)"
timeout 300 python -m pytest tests/test_lines_gpu.py -x -q -m gpu 2>&1 | tail -1
run --steps 20 --warmup 5
run --steps 64
} 2>&1 | tee gpurun_out/r02_s20_sweep.log
