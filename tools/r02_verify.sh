#!/bin/bash
TAG=${1:-r02v}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 150 2>&1 | tail -3 | tee gpurun_out/${TAG}_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-latency > gpurun_out/${TAG}_b64.json 2> gpurun_out/${TAG}_b64.err; python tools/benchline.py ${TAG}_64 < gpurun_out/${TAG}_b64.json
timeout 300 python bench.py --no-cpu-baseline --no-latency --steps 20 --warmup 5 > gpurun_out/${TAG}_b20.json 2> gpurun_out/${TAG}_b20.err; python tools/benchline.py ${TAG}_20 < gpurun_out/${TAG}_b20.json
python -c "import json; d=json.load(open('gpurun_out/${TAG}_b20.json')); print(d['e2e']['api'][:140]); print(d['e2e']['by_api'])"
