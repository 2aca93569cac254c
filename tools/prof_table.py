"""Print bench.py's per-shape event table (--profile-json) sorted by time.  usage: python tools/prof_table.py file.json [n]"""
import json
import sys

d = json.load(open(sys.argv[1]))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print(f"eager step {d['eager_step_ms'] * 1e3:.0f} us, graph step {d['graph_step_ms'] * 1e3:.0f} us")
for r in d["per_kernel"]:
    print(f"  {r['kernel']:28s} n={r['n']:2d} us/launch={r['us_per_launch']:6.1f} share={r['share']:.3f} hbm_frac={r['hbm_frac']:.2f}")
for r in sorted(d["kernels"], key=lambda r: -r["ms"])[:top]:
    print(f"{r['name']:22s} {r['shape']:40s} n={r['n']} us={1e3 * r['ms'] / r['n']:6.1f} tot={1e3 * r['ms']:6.1f} MB={r['bytes'] / 1e6:6.1f} hbm={r['hbm_frac']:.2f}")
