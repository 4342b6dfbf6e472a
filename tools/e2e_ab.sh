#!/bin/bash
# A/B of the two host paths on one box: stream-ordered submit_host vs host-scheduled slots
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_frontend_gpu.py -x -q -m gpu 2>&1 | tail -3
for m in 0 1 0 1; do
  PLSLAM_E2E_SLOTS=$m timeout 100 python bench.py --no-cpu-baseline --steps 32 > gpurun_out/e2e_ab_$m.json 2> gpurun_out/e2e_ab_$m.err
  python tools/benchline.py "slots=$m" < gpurun_out/e2e_ab_$m.json
done
