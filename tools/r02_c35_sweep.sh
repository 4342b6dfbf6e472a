#!/bin/bash
# C3 (1280x720) and C5 (3840x2160): throughput against the number of batches in flight
for cfg in "c3 4 16" "c3 12 36" "c3 24 48" "c5 2 6" "c5 6 12" "c5 12 24"; do
  set -- $cfg
  timeout 600 python bench.py --workload $1 --depth $2 --steps $3 --no-cpu-baseline --no-latency 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 depth $2: value %.1f e2e %.1f ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d['roofline']['stages_ms'].get('lsd_grow'))"
  nvidia-smi --query-gpu=memory.used --format=csv,noheader
done
