"""Turn an ncu launch list (--metrics gpu__time_duration.sum --csv) of `bench.py --launch-list --no-graph`
into a per-kernel table of the LAST step: share of the step, launches, mean / total duration.

    python tools/launch_table.py gpurun_out/launches.csv [launches_per_step] > profiles/launches_rNN.md
"""
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
per_step = int(sys.argv[2]) if len(sys.argv) > 2 else 124
step = rows[-per_step:]


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("mgdt::", "")
    return name


agg = {}
for r in step:
    a = agg.setdefault(short(r[4]), [0.0, 0])
    a[0] += float(r[14]) / 1000.0
    a[1] += 1
tot = sum(a[0] for a in agg.values())
print(f"# ncu launch list, last step: {len(step)} launches, {tot:.1f} us serialised (cold-cache per-launch times)\n")
print("| kernel | launches | total us | mean us | share |")
print("|---|---:|---:|---:|---:|")
for n, (t, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"| `{n}` | {c} | {t:.1f} | {t / c:.1f} | {100 * t / tot:.1f}% |")
print("\n## launches in order\n")
print("| # | kernel | grid | block | us |")
print("|---:|---|---|---|---:|")
for i, r in enumerate(step):
    print(f"| {i} | `{short(r[4])}` | {r[8]} | {r[7]} | {float(r[14]) / 1000.0:.1f} |")
