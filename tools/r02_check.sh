#!/bin/bash
# round-2 check on one box: GPU tests, default bench line, launch list under ncu (profiles/r02_*), smoke
TAG=${1:-r02a}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/${TAG}_tests.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python tools/benchline.py $TAG -v < gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --depth 1 --no-cpu-baseline --no-latency > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv | head -40
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
# full captures of the dominant kernel and of the changed second-tier kernels (one launch each)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_lsd_grow|k_fast|k_lbd|k_lsd_nfa$|k_lsd_grad" -s 10 -c 5 -o gpurun_out/${TAG}_full \
  python bench.py --steps 2 --warmup 1 --depth 1 --no-cpu-baseline --no-latency > gpurun_out/${TAG}_full_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_full_ncu.log
