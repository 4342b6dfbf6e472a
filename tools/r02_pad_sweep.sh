#!/bin/bash
# residency cap of k_lsd_grow (extra dynamic shared memory per CTA) against pipelined throughput
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lines_gpu.py tests/test_frontend_gpu.py -x -q -m gpu 2>&1 | tail -2
for pad in 0 5000 10000 17000 29000; do
  export PLSLAM_GROW_PAD=$pad
  echo "#### PAD=$pad"
  timeout 300 python bench.py --no-cpu-baseline --no-latency --steps 48 > gpurun_out/pad.json 2> gpurun_out/pad.err
  python tools/benchline.py pad < gpurun_out/pad.json
done 2>&1 | tee gpurun_out/r02_pad_sweep.log
