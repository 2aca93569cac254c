#!/bin/bash
# Reduced form of capture_profiles.sh whose output stays under gpurun's 64 MiB pull limit: launch list, DRAM traffic of
# the convolution-class kernels, and ONE `--set full` capture of the kernels added late in round 2 (stem_mma,
# conv3x3_warp, tood_cls), exported to CSV on the box (the .ncu-rep files are deleted there).
set -u
R=${1:-r02}
mkdir -p gpurun_out
python bench.py --launch-list --no-graph --steps 1 --warmup 3 > gpurun_out/ll_plain_$R.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --launch-list --no-graph --steps 1 --warmup 3 > gpurun_out/ll_ncu_$R.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:"conv_umma2|conv1x1_tma|conv3x3_tma|conv3x3_warp|stem_mma" --clock-control none \
    -c 1600 --csv --log-file gpurun_out/traffic_$R.csv python bench.py --launch-list --no-graph --steps 1 --warmup 3 \
    > gpurun_out/traffic_ncu_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"stem_mma|conv3x3_warp|tood_cls" --launch-skip ${SKIP:-45} -c 15 -f \
    -o /tmp/newk_$R python bench.py --launch-list --no-graph --steps 1 --warmup 3 > gpurun_out/newk_ncu_$R.log 2>&1
ncu -i /tmp/newk_$R.ncu-rep --page raw --csv > gpurun_out/newk_raw_$R.csv 2>/dev/null
ncu -i /tmp/newk_$R.ncu-rep --page source --csv --print-source sass --kernel-id :::1 > gpurun_out/newk_src_stem_$R.csv 2>/dev/null
gzip -f gpurun_out/launches_$R.csv gpurun_out/traffic_$R.csv gpurun_out/newk_src_stem_$R.csv
ls -la gpurun_out | tail -8
