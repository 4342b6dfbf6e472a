#!/bin/bash
# round-end check on one box: GPU tests, the default bench line, smoke
mkdir -p gpurun_out
timeout 200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 200 python bench.py > gpurun_out/final4_bench.json 2> gpurun_out/final4_bench.err
tail -3 gpurun_out/final4_bench.err
python tools/benchline.py final < gpurun_out/final4_bench.json
python -c "import json; d=json.load(open('gpurun_out/final4_bench.json')); print(d['e2e']['by_api'], d['steps'])"
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
