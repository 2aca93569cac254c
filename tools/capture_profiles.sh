#!/bin/bash
# Run ON THE GPU BOX (through gpurun) from the repo root: regenerates the raw ncu material the files under
# profiles/ are derived from.  Each ncu command runs only after the same program has exited 0 without ncu.
set -u
R=${1:-r01}
mkdir -p gpurun_out
python bench.py --launch-list --no-graph --steps 1 --warmup 3 > gpurun_out/ll_plain_$R.log 2>&1 || exit 1
# (1) launch list: per-launch durations of every kernel of the step
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --launch-list --no-graph --steps 1 --warmup 3 > gpurun_out/ll_ncu_$R.log 2>&1
# (2) measured DRAM traffic of the dominant kernel, every launch of one step
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:"conv_umma2|conv1x1_tma|conv3x3_tma|conv3x3_warp|stem_mma" --clock-control none \
    -c 1600 --csv --log-file gpurun_out/traffic_$R.csv python bench.py --launch-list --no-graph --steps 1 --warmup 3 \
    > gpurun_out/traffic_ncu_$R.log 2>&1
# (3) one full capture of representative launches of the dominant kernel (source-level stalls, pipe utilisation)
python tools/ncu_conv.py c64_256 c32_3x3 c96_384 c16_32_s2 > gpurun_out/ncu_conv_plain_$R.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"conv_umma2|conv1x1_tma|conv3x3_tma" --launch-skip 0 -c 8 -f \
    -o gpurun_out/conv_umma2_full_$R python tools/ncu_conv.py c64_256 c32_3x3 c96_384 c16_32_s2 > gpurun_out/ncu_conv_full_$R.log 2>&1
ls -la gpurun_out | tail -8
# (4) full capture of the non-conv (memory-bound) kernels of the step -> python tools/ncu_membound.py ... profiles/membound_$R.md
ncu --set full --clock-control none -k regex:"affine_act|inject2x|mspa_front|avgpool|bilinear2x|dwconv7|decode_|sppf_pool|resample_kernel|chan_stats|nms_scan|stem_mma|conv3x3_warp|tood_cls" \
    --launch-skip ${SKIP:-102} -c 34 -f -o gpurun_out/membound_$R python bench.py --launch-list --no-graph --steps 1 --warmup 3 > gpurun_out/membound_ncu_$R.log 2>&1
