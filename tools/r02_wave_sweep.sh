#!/bin/bash
# e2e leg of bench.py against the wave size of plslam_frontend_submit_host_wave (waves rotate over the slots)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_frontend_gpu.py -x -q -m gpu 2>&1 | tail -2
for cfg in "64 32" "64 16" "64 8" "64 4" "20 5" "20 20"; do
  set -- $cfg
  echo "#### steps=$1 wave=$2"
  timeout 300 python bench.py --no-cpu-baseline --no-latency --steps $1 --wave $2 > gpurun_out/wave.json 2> gpurun_out/wave.err || tail -5 gpurun_out/wave.err
  python tools/benchline.py wave < gpurun_out/wave.json
  python -c "import json; d=json.load(open('gpurun_out/wave.json')); print(d['e2e']['by_api'])"
done 2>&1 | tee gpurun_out/r02_wave_sweep.log
