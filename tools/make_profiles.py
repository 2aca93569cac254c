"""Derive the committed summaries under profiles/ from the raw ncu material in gpurun_out/ (tools/capture_profiles.sh).

    python tools/make_profiles.py r01 <conv launches per step> <launches per step>
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
CONV_PER_STEP = int(sys.argv[2]) if len(sys.argv) > 2 else 57
PER_STEP = int(sys.argv[3]) if len(sys.argv) > 3 else 113
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("mgdt::", "")


# ---- (1) launch list
rows = [r for r in csv.reader(open(os.path.join(G, f"launches_{R}.csv"))) if len(r) > 14 and r[0].isdigit()]
names = [short(r[4]) for r in rows]
stems = [i for i, n in enumerate(names) if n.startswith("conv_umma2_kernel<0, 3") or n.startswith("stem_mma_kernel")]   # layer 0 = first launch of a step
start = stems[-1]
step = rows[start:start + PER_STEP]
agg = {}
for r in step:
    a = agg.setdefault(short(r[4]), [0.0, 0])
    a[0] += float(r[14]) / 1000.0
    a[1] += 1
tot = sum(a[0] for a in agg.values())
fam = {}
for n, (t, c) in agg.items():
    f = re.sub(r"<.*", "", n)
    a = fam.setdefault(f, [0.0, 0])
    a[0] += t
    a[1] += c
with open(os.path.join(P, f"launches_{R}.md"), "w") as f:
    f.write(f"# ncu launch list -- last step of `bench.py --launch-list --no-graph --steps 1 --warmup 3` (B=32, bf16)\n\n"
            f"`ncu --metrics gpu__time_duration.sum --clock-control none`; {len(step)} launches, {tot:.1f} us serialised.\n"
            "Per-launch times under ncu are cold-cache and serialised (no overlap between the two in-flight batches, no\n"
            "programmatic dependent launch), so the SHARES are what compares with bench.py's CUDA-event profile.\n\n"
            "## by kernel function\n\n| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n")
    for n, (t, c) in sorted(fam.items(), key=lambda x: -x[1][0]):
        f.write(f"| `{n}` | {c} | {t:.1f} | {t / c:.1f} | {100 * t / tot:.1f}% |\n")
    f.write("\n## by template instantiation (`conv_umma2_kernel<MODE, LOADER, SPLIT>`: MODE 0 = 1x1, 1 = 3x3 s1, 2 = 3x3 s2; "
            "LOADER 0 = cp.async, 1 = cp.async + in-place transform, 2 = DCN sampler, 3 = uint8 stem; "
            "SPLIT 0/1/2 = 3/16, 7/12, 11/8 producer/epilogue warps)\n\n| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n")
    for n, (t, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
        f.write(f"| `{n}` | {c} | {t:.1f} | {t / c:.1f} | {100 * t / tot:.1f}% |\n")
    f.write("\n## launches in order\n\n| # | kernel | grid | block | us |\n|---:|---|---|---|---:|\n")
    for i, r in enumerate(step):
        f.write(f"| {i} | `{short(r[4])}` | {r[8]} | {r[7]} | {float(r[14]) / 1000.0:.1f} |\n")
conv_fams = [k for k in fam if k.startswith("conv")]
print("launch list:", len(step), "launches", f"{tot:.1f} us;", ", ".join(f"{k} {100 * fam[k][0] / tot:.1f}%" for k in conv_fams))

# ---- (2) traffic
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_traffic.py"), os.path.join(G, f"traffic_{R}.csv"),
                      "ALL" if R != "r01" else "conv_umma2_kernel", str(CONV_PER_STEP), "32", "mspa_c2f_gd_tood_yolov8n"], capture_output=True, text=True)
open(os.path.join(P, f"traffic_{R}.json"), "w").write(out.stdout)
print(out.stdout)

# ---- (3) full capture: per-launch key metrics + top stall lines
rep = os.path.join(G, f"conv_umma2_full_{R}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units = rr[0], rr[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tc.sum", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
TC_WANT = ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tc.sum", "smsp__inst_executed_pipe_uniform.sum")
tc_cols = [(h, i) for i, h in enumerate(hdr) if h in TC_WANT]
cases = ["c32_3x3 (32->32 3x3 @80^2, B=32)", "c32_3x3 (2nd launch, warm)", "c96_384 (96->384 1x1 @40^2)", "c96_384 (2nd launch, warm)",
         "c64_256 (64->256 1x1 @80^2)", "c64_256 (2nd launch, warm)"]
if R != "r01":
    cases = ["c32_3x3 (32->32 3x3 @80^2, B=32: conv3x3_tma_kernel)", "c32_3x3 (2nd launch, warm)",
             "c16_32_s2 (16->32 3x3 stride 2 @320^2: conv_umma2_kernel)", "c16_32_s2 (2nd launch, warm)",
             "c96_384 (96->384 1x1 @40^2: conv1x1_tma_kernel without statistics)", "c96_384 (2nd launch, warm)",
             "c64_256 (64->256 1x1 @80^2: conv1x1_tma_kernel)", "c64_256 (2nd launch, warm)"]
with open(os.path.join(P, f"conv_umma2_full_{R}.md"), "w") as f:
    f.write(f"# `ncu --set full --clock-control none --import-source on` of the tcgen05 conv kernels -- tools/ncu_conv.py ({', '.join(c.split()[0] for c in cases[::2])})\n\n"
            "Two launches per shape (the second is warm).  Raw report: gpurun_out/ (scratch, not committed).\n\n")
    for k, r in enumerate(rr[2:]):
        f.write(f"## launch {k}: {cases[k] if k < len(cases) else ''}\n\n| metric | value | unit |\n|---|---:|---|\n")
        for w, i in idx:
            f.write(f"| {w} | {r[i][:60]} | {units[i]} |\n")
        for h, i in tc_cols:
            if (h, i) not in [(a, b) for a, b in idx] and r[i] not in ("", "0", "n/a"):
                f.write(f"| {h} | {r[i][:40]} | {units[i]} |\n")
        f.write("\n")
    # stall summary of the warm c64_256 and c32_3x3 launches
    for kid, label in (((2, "c32_3x3, warm launch"), (6, "c64_256, warm launch")) if R == "r01" else
                       ((2, "c32_3x3 (conv3x3_tma_kernel), warm launch"), (4, "c16_32_s2 (conv_umma2_kernel), warm launch"),
                        (8, "c64_256 (conv1x1_tma_kernel), warm launch"))):
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-id", f":::{kid}"],
                             capture_output=True, text=True).stdout
        sr = list(csv.reader(io.StringIO(src)))
        try:
            hi = next(i for i, r in enumerate(sr) if "Source" in r and "# Samples" in r)
        except StopIteration:
            continue
        h2 = sr[hi]
        i_src, i_s = h2.index("Source"), h2.index("# Samples")
        stalls = [(h, j) for j, h in enumerate(h2) if h.startswith("stall_") and "Not Issued" not in h]
        data, seen = [], set()
        for r in sr[hi + 1:]:   # (the source page lists every instruction twice: dedupe on the address column)
            try:
                if r[0] in seen:
                    continue
                seen.add(r[0])
                data.append((int(r[i_s]), r))
            except Exception:
                pass
        tot_s = sum(n for n, _ in data) or 1
        sagg = {}
        for n, r in data:
            for h, j in stalls:
                sagg[h] = sagg.get(h, 0) + int(r[j] or 0)
        f.write(f"## warp-stall sampling, {label}: {tot_s} samples over {len(data)} SASS instructions\n\n")
        f.write("stall totals: " + ", ".join(f"{h[6:]} {100 * v // tot_s}%" for h, v in sorted(sagg.items(), key=lambda x: -x[1])[:8]) + "\n\n")
        f.write("| samples | share | SASS | top stalls |\n|---:|---:|---|---|\n")
        for n, r in sorted(data, key=lambda x: -x[0])[:14]:
            st = sorted(((int(r[j] or 0), h[6:]) for h, j in stalls), reverse=True)[:2]
            f.write(f"| {n} | {100 * n / tot_s:.1f}% | `{r[i_src][:60].strip()}` | {st[0][1]} {st[0][0]}, {st[1][1]} {st[1][0]} |\n")
        f.write("\n")
print("wrote", os.path.join(P, f"conv_umma2_full_{R}.md"))
