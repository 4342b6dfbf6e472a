#!/bin/bash
# end-of-round-2 evidence on one box: GPU tests, bench lines (default and the driver's --steps 20 --warmup 5), launch list and
# full captures under ncu, smoke, C3 / C5 lines
TAG=${1:-r02z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 150 2>&1 | tail -4 | tee gpurun_out/${TAG}_tests.log
timeout 500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -2 gpurun_out/${TAG}_bench.err
python tools/benchline.py $TAG -v < gpurun_out/${TAG}_bench.json
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_s20.json 2> gpurun_out/${TAG}_bench_s20.err
python tools/benchline.py ${TAG}_s20 < gpurun_out/${TAG}_bench_s20.json
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
export PLSLAM_GROW_MODE=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --depth 1 --no-cpu-baseline --no-latency > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv | head -40
if [ -z "$SKIP_FULL" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_lsd_grow|k_lsd_nfa$|k_lsd_grad|k_lbd|k_lsd_scatter|k_lsd_rowhist" -s 12 -c 6 -o gpurun_out/${TAG}_full \
  python bench.py --steps 2 --warmup 1 --depth 1 --no-cpu-baseline --no-latency > gpurun_out/${TAG}_full_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_full_ncu.log
fi
unset PLSLAM_GROW_MODE
for w in c3 c4 c5; do
  timeout 400 python bench.py --workload $w --no-cpu-baseline --no-latency > gpurun_out/${TAG}_$w.json 2> gpurun_out/${TAG}_$w.err
  python tools/benchline.py ${TAG}_$w < gpurun_out/${TAG}_$w.json
done
