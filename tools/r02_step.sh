#!/bin/bash
# quick check after a kernel change: line/LBD parity tests, stage times of one batch, throughput with 16 batches in flight
TAG=${1:-r02s}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lines_gpu.py tests/test_lbd_kat.py tests/test_frontend_gpu.py tests/test_orb_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --no-latency > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python tools/benchline.py $TAG -v < gpurun_out/${TAG}_bench.json
