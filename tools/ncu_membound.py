"""profiles/membound_rNN.md from an `ncu --set full` report of the non-conv kernels of one eager step.

Capture (on the GPU box, after the same command has exited 0 without ncu):
    ncu --set full --clock-control none -k regex:"affine_act|inject2x|mspa_front|avgpool|bilinear2x|dwconv7|decode_staged|sppf_pool|resample_kernel|chan_stats|nms_scan" \
        --launch-skip 120 -c 30 -f -o gpurun_out/membound_r01 python bench.py --launch-list --no-graph --steps 1 --warmup 3
Then here:
    python tools/ncu_membound.py gpurun_out/membound_r01.ncu-rep profiles/membound_r01.md
"""
import csv
import io
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units, rows = rr[0], rr[1], rr[2:]
col = {h: i for i, h in enumerate(hdr)}


def f(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


tmul = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0)
bmul = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
rmul = bmul.get(units[col["dram__bytes_read.sum"]], 1.0)
wmul = bmul.get(units[col["dram__bytes_write.sum"]], 1.0)
PEAK = 6555.5   # measured copy bandwidth, MEASURED_PEAKS.json
lines = [
    "# `ncu --set full --clock-control none` of the memory-bound kernels -- last eager step of `bench.py --launch-list --no-graph` (B = 32, bf16)\n",
    "One launch per line (cold cache, serialised).  GB/s = (DRAM bytes read + written) / duration; `% of peak` is against the",
    f"measured copy bandwidth {PEAK} GB/s (MEASURED_PEAKS.json).  DRAM WRITE bytes are ~0 for every kernel whose output is",
    "<= ~30 MB: the 126 MB L2 absorbs the stores and writes them back after the kernel, so for those launches the DRAM figure",
    "is the read side only and the kernel is bound by L2 / latency / launch shape, not by HBM.\n",
    "| kernel | grid | us | DRAM read MB | DRAM write MB | GB/s | % of peak | L2 hit % | SM % | regs | warps active % |",
    "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
for r in rows:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("mgdt::", "")
    us = f(r[col["gpu__time_duration.sum"]]) * tmul
    rd, wr = f(r[col["dram__bytes_read.sum"]]) * rmul, f(r[col["dram__bytes_write.sum"]]) * wmul
    gbs = (rd + wr) / us * 1e3
    lines.append(f"| `{name}` | {r[col['Grid Size']]} | {us:.1f} | {rd:.1f} | {wr:.1f} | {gbs:.0f} | {100 * gbs / PEAK:.0f}% | "
                 f"{f(r[col['lts__t_sector_hit_rate.pct']]):.0f} | {f(r[col['sm__throughput.avg.pct_of_peak_sustained_elapsed']]):.0f} | "
                 f"{r[col['launch__registers_per_thread']]} | {f(r[col['sm__warps_active.avg.pct_of_peak_sustained_active']]):.0f} |")
open(out, "w").write("\n".join(lines) + "\n")
print(out, len(rows), "launches")
