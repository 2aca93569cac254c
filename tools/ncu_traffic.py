"""profiles/traffic_rNN.json from an ncu CSV of `bench.py --launch-list --no-graph` restricted to one kernel:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:conv_umma2 \
        --clock-control none --csv --log-file gpurun_out/traffic.csv python bench.py --launch-list --no-graph --steps 1 --warmup 3
    python tools/ncu_traffic.py gpurun_out/traffic.csv conv_umma2_kernel 63 32 mspa_c2f_gd_tood_yolov8n > profiles/traffic_r01.json

Takes the LAST `n` launches of the kernel (= one step) and reports measured DRAM bytes per launch.
"""
import csv
import json
import sys

path, kernel, n, batch, workload = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
if kernel == "ALL":   # one entry per conv kernel function: {"kernels": {name: {...}}}; n = launches of ALL of them in one step
    import re
    import subprocess
    allrows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    ids_all = sorted({int(r[0]) for r in allrows})[-n:]
    names = {}
    for r in allrows:
        if int(r[0]) in ids_all:
            names.setdefault(re.sub(r"<.*", "", r[4].replace("void ", "").replace("mgdt::", "")), set()).add(int(r[0]))
    out = {}
    for k, ids in names.items():
        sub = subprocess.run([sys.executable, __file__, path, k, str(len(ids)), str(batch), workload], capture_output=True, text=True)
        out[k] = json.loads(sub.stdout)
    print(json.dumps({"kernels": out}, indent=1))
    sys.exit(0)
rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit() and kernel in r[4]]
by_id = {}
for r in rows:
    d = by_id.setdefault(int(r[0]), {})
    val = float(r[14].replace(",", ""))
    unit = r[13]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}.get(unit, 1.0)
    d[r[12]] = val * scale
ids = sorted(by_id)[-n:]
rd = sum(by_id[i].get("dram__bytes_read.sum", 0.0) for i in ids)
wr = sum(by_id[i].get("dram__bytes_write.sum", 0.0) for i in ids)
ns = sum(by_id[i].get("gpu__time_duration.sum", 0.0) for i in ids)
print(json.dumps({"kernel": kernel, "workload": workload, "batch": batch, "launches": len(ids),
                  "dram_read_bytes_per_step": rd, "dram_write_bytes_per_step": wr,
                  "dram_bytes_per_launch": (rd + wr) / max(len(ids), 1),
                  "ncu_duration_us_per_launch": ns / 1e3 / max(len(ids), 1),
                  "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
                            "--clock-control none, last step of bench.py --launch-list --no-graph"}, indent=1))
