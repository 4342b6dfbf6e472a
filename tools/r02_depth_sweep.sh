#!/bin/bash
# frames in flight held at 4096: batch x depth trade-off, and the line branch alone with / without its tail
for cfg in "256 16 64" "512 8 32" "1024 4 16" "2048 2 8" "256 32 96" "512 16 48"; do
  set -- $cfg
  python bench.py --batch $1 --depth $2 --steps $3 --no-cpu-baseline --no-latency 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('batch $1 depth $2: value %.0f e2e %.0f ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
echo "lines only, 16 in flight:"; python tools/prof_lines_conc.py 16
echo "lines only, stop after grow, 16 in flight:"; PLSLAM_DEBUG_STOP_AFTER_GROW=1 python tools/prof_lines_conc.py 16
