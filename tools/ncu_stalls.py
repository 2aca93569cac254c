"""Top SASS instructions by warp-stall samples of one kernel of an `ncu --set full --import-source on` report.

    python tools/ncu_stalls.py report.ncu-rep [kernel id, default 1] [rows, default 25]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else "1"
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr = rr[0]
for name in ("Kernel Name", "gpu__time_duration.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
             "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
             "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
             "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread"):
    if name in hdr:
        i = hdr.index(name)
        print(f"{name:70s}", " | ".join(r[i][:40] for r in rr[2:]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-id", f":::{kid}"],
                     capture_output=True, text=True).stdout
sr = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(sr) if "Source" in r and "# Samples" in r)
h2 = sr[hi]
i_src, i_s = h2.index("Source"), h2.index("# Samples")
stalls = [(h, j) for j, h in enumerate(h2) if h.startswith("stall_") and "Not Issued" not in h]
data, seen = [], set()
for r in sr[hi + 1:]:
    try:
        if r[0] in seen:
            continue
        seen.add(r[0])
        data.append((int(r[i_s]), r))
    except Exception:
        pass
tot = sum(n for n, _ in data) or 1
agg = {}
for n, r in data:
    for h, j in stalls:
        agg[h] = agg.get(h, 0) + int(r[j] or 0)
print(f"\n{tot} samples over {len(data)} instructions; stall totals: " +
      ", ".join(f"{h[6:]} {100 * v // tot}%" for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for n, r in sorted(data, key=lambda x: -x[0])[:rows]:
    st = sorted(((int(r[j] or 0), h[6:]) for h, j in stalls), reverse=True)[:2]
    print(f"{n:6d} {100 * n / tot:5.1f}%  {r[i_src][:70].strip():70s}  {st[0][1]} {st[0][0]}, {st[1][1]} {st[1][0]}")
