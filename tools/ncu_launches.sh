#!/bin/bash
# Run ON THE GPU BOX (through gpurun): ncu launch list of one eager step (per-launch durations), tag = $1
set -u
R=${1:-x}
python bench.py --launch-list --no-graph --steps 1 --warmup 3 > gpurun_out/ll_plain_$R.log 2>&1 || { tail -5 gpurun_out/ll_plain_$R.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --launch-list --no-graph --steps 1 --warmup 3 > gpurun_out/ll_ncu_$R.log 2>&1
tail -2 gpurun_out/ll_plain_$R.log | cut -c1-200
