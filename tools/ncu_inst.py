"""Executed-instruction profile of an ncu report (source page): per-opcode totals and the hottest SASS lines.

    python tools/ncu_inst.py REPORT.ncu-rep [N]
"""
import csv, io, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
i_src = hdr.index("Source")
i_ex = next(j for j, h in enumerate(hdr) if h.strip() in ("Instructions Executed", "# Instructions Executed", "Warp Instructions Executed"))
tot, ops, lines = 0, {}, []
for k, r in enumerate(rows[hi + 1:]):
    try:
        n = int(r[i_ex])
    except Exception:
        continue
    src = r[i_src].strip()
    op = src.split()[1] if src.startswith("@") and len(src.split()) > 1 else src.split()[0] if src else "?"
    op = op.split(".")[0]
    tot += n
    ops[op] = ops.get(op, 0) + n
    lines.append((n, k, src))
print("warp instructions executed:", tot)
for op, n in sorted(ops.items(), key=lambda x: -x[1])[:25]:
    print(f"  {op:12s} {n:10d} {n / tot * 100:5.1f}%")
# cumulative executed count by line index (to see which region of the kernel the instructions come from)
lines.sort(key=lambda x: x[1])
acc, marks = 0, []
for n, k, src in lines:
    acc += n
    marks.append((k, acc))
step = max(1, len(marks) // 40)
print("cumulative share by SASS line index:")
print("  " + " ".join(f"{k}:{a / tot * 100:.0f}%" for k, a in marks[::step]))
print("hottest lines:")
for n, k, src in sorted(lines, key=lambda x: -x[0])[:topn]:
    print(f"  {n:9d} #{k:5d} {src[:90]}")
