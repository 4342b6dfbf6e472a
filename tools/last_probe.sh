#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_frontend_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 120 python bench.py > gpurun_out/final5_bench.json 2> gpurun_out/final5_bench.err
tail -2 gpurun_out/final5_bench.err
python tools/benchline.py final5 < gpurun_out/final5_bench.json
python -c "import json; d=json.load(open('gpurun_out/final5_bench.json')); print(d['e2e']['by_api'])"
