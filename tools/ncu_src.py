"""Summarise an ncu report's source page: top stalled SASS instructions with their dominant stall reasons.

    python tools/ncu_src.py REPORT.ncu-rep [N]
"""
import csv, subprocess, sys, io
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
i_src, i_s = hdr.index("Source"), hdr.index("# Samples")
stalls = [(h, j) for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for k, r in enumerate(rows[hi + 1:]):
    try:
        n = int(r[i_s])
    except Exception:
        continue
    data.append((n, k, r))
tot = sum(n for n, _, _ in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for n, k, r in data:
    for h, j in stalls:
        agg[h] = agg.get(h, 0) + int(r[j] or 0)
print("stall totals:", ", ".join(f"{h[6:]}={v*100//max(tot,1)}%" for h, v in sorted(agg.items(), key=lambda x: -x[1])[:10]))
for n, k, r in sorted(data, key=lambda x: -x[0])[:topn]:
    st = sorted(((int(r[j] or 0), h[6:]) for h, j in stalls), reverse=True)[:2]
    print(f"{n:6d} {n/tot*100:5.1f}% #{k:5d} {r[i_src][:70]:70s} {st}")
