#!/bin/bash
# e2e leg of the driver's 20-step invocation against the sizes of the first waves
mkdir -p gpurun_out
export PLSLAM_ORB_AFTER=1
for r in "2,4,14" "4,16" "1,3,6,10" "2,2,4,4,8"; do
  echo "#### ramp=$r"
  timeout 300 python bench.py --no-cpu-baseline --no-latency --steps 20 --warmup 5 --wave-ramp $r > gpurun_out/ramp.json 2> gpurun_out/ramp.err || tail -3 gpurun_out/ramp.err
  python tools/benchline.py ramp < gpurun_out/ramp.json
  python -c "import json; d=json.load(open('gpurun_out/ramp.json')); print(d['e2e']['by_api'])"
done 2>&1 | tee gpurun_out/r02_ramp.log
