"""profiles/newkernels_rNN.md from the raw-page CSV of the `--set full` capture in tools/capture_profiles_small.sh.

    python tools/ncu_newk.py gpurun_out/newk_raw_r02.csv profiles/newkernels_r02.md
"""
import csv
import re
import sys

src, out = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def f(r, name):
    try:
        return float(r[col[name]].replace(",", ""))
    except (ValueError, KeyError):
        return float("nan")


PEAK = 6555.5
lines = ["# `ncu --set full --clock-control none` of the kernels added late in round 2 -- last eager step of `bench.py --launch-list --no-graph` (B = 32, bf16)\n",
         "One launch per line (cold cache, serialised; durations under ncu).  `HMMA %` = sm__pipe_tensor_subpipe_hmma_cycles_active (warp-level",
         "mma.sync; measured peak 1024 MAC / clk / SM, `tools/ubench/hmma_rate.cu`).  DRAM bytes are the read side for maps that fit the 126 MB L2.\n",
         "| kernel | grid | us | DRAM read MB | DRAM write MB | GB/s | % of HBM peak | regs | warps active % | warp instr | IPC / SM | HMMA % | L2 hit % |",
         "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
for r in data:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("mgdt::", "")
    us, rd, wr = f(r, "gpu__time_duration.sum"), f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")
    gbs = (rd + wr) / us * 1e3
    lines.append(f"| `{name}` | {r[col['Grid Size']]} | {us:.1f} | {rd:.1f} | {wr:.1f} | {gbs:.0f} | {100 * gbs / PEAK:.0f}% | "
                 f"{r[col['launch__registers_per_thread']]} | {f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | "
                 f"{f(r, 'smsp__inst_executed.sum') / 1e6:.2f} M | {f(r, 'sm__inst_executed.avg.per_cycle_elapsed'):.2f} | "
                 f"{f(r, 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active'):.0f} | {f(r, 'lts__t_sector_hit_rate.pct'):.0f} |")
open(out, "w").write("\n".join(lines) + "\n")
print(out, len(data), "launches")
