#!/usr/bin/env python
"""bench.py -- images/s @640x640 for forward + DFL decode + NMS (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--batch B]

One process per GPU (the driver launches N>1 under torch.distributed.run); the batch shards over
the ranks with no data-path collective (weak scaling: `--batch` images per GPU per step).

JSON line (rank 0):
  value         whole-job images/s with the uint8 input batches already resident in HBM
  e2e           the same through Engine.submit()/collect(): pinned-host uint8 batch -> H2D ->
                preprocess -> forward -> decode -> NMS -> D2H of the packed detections, all inside
                the timed region (two batches in flight so copies overlap kernels)
  roofline      the dominant kernel (largest share of device time), algorithmic bytes / CUDA-event time
  cpu_baseline  the oracle port (torch CPU fp32, all host threads) on a bounded sample, rank 0 / N=1

`--impl reference` times that CPU port alone (the reference is pure Python/PyTorch and cannot
travel to the GPU box; oracle/mgdt_oracle.py is its restatement, pinned to the live reference by
tests/test_oracle_golden.py).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (cfg, nc, class-predictor bias).  Random-init class logits never pass `conf`
    # (head.py:186,568), which would make NMS trivial; the bias is calibrated (oracle, 4 synthetic
    # images, seed 1) so that 5 % of the anchors pass conf 0.25.
    "mspa_c2f_gd_tood_yolov8n": ("mspa_c2f_gd_tood_yolov8n.yaml", 2, -1.238),  # BASELINE.json configs[3]: full MGDT
    "mspa_c2f_gd_yolov8n": ("mspa_c2f_gd_yolov8n.yaml", 80, -18.701),
    "mspa_c2f_yolov8n": ("mspa_c2f_yolov8n.yaml", 80, -1.703),
    "yolov8n": ("yolov8n.yaml", 80, -33.456),
    # other width scales (the TOODHead YAMLs fix the head width and exist at scale n only); the bias is
    # the n-scale one, so the share of anchors that pass conf differs
    "mspa_c2f_gd_yolov8s": ("mspa_c2f_gd_yolov8s.yaml", 80, -18.701),
    "mspa_c2f_yolov8m": ("mspa_c2f_yolov8m.yaml", 80, -1.703),
}
CONF, IOU, MAX_DET = 0.25, 0.7, 300


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mspa_c2f_gd_tood_yolov8n", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--slots", type=int, default=3, help="batches in flight (engine slots, one CUDA stream each; measured 2 / 3 / 4 -> 17.6k / 18.2k / 18.4k images/s, e2e 15.8k / 16.8k / 16.5k)")
    ap.add_argument("--profile-json", default="", help="write the per-launch table here")
    ap.add_argument("--launch-list", action="store_true",
                    help="eager warm-up + timed steps only (no e2e / per-launch / CPU legs): the run to put under ncu")
    return ap.parse_args()


# ------------------------------------------------------------------------------------ helpers
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, cmax = float(f[1]), float(f[2])
            except ValueError:
                continue
            mx = max(mx, cmax)
            if t0 - 0.05 <= t <= t1 + 0.15:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: use the nearest samples
            sm = [float(l.split(",")[1]) for _, l in self.rows[-3:] if l.count(",") >= 8] or [0.0]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p["hbm_gbs"], p["bf16_tflops"], p.get("bf16_tflops_sustained", p["bf16_tflops"]), "measured"
    except Exception:
        return 6650.0, 1590.0, 1400.0, "fallback"


def make_u8(batch, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (batch, 3, 640, 640), dtype=torch.uint8, generator=g)


# ------------------------------------------------------------------------------------ CPU arm
def cpu_reference(workload, sample_images, repeats):
    """The reference's CPU path restated (oracle port): fused-BN fp32 forward + decode + NMS with all
    host threads.  Returns (images/s, seconds, threads)."""
    import torch
    from mgdt_yolo_b200.synth import raise_cls_bias, synth_state_dict
    from mgdt_yolo_b200.tasks import DetectionModel
    from oracle import mgdt_oracle as O
    cfg, nc, cls_bias = WORKLOADS[workload]
    threads = host_threads()
    torch.set_num_threads(threads)
    tmpl = DetectionModel(cfg, nc=nc, verbose=False).state_dict()  # shapes only (CPU, no arithmetic)
    sd = O.fold_bn(raise_cls_bias(synth_state_dict(tmpl, seed=1), cls_bias))  # AutoBackend fuses (autobackend.py:97)
    x = make_u8(sample_images, 0).float() / 255
    times = []
    with torch.inference_mode():
        for it in range(repeats + 1):
            t0 = time.perf_counter()
            y, _, _ = O.forward(cfg, sd, x, nc=nc, dcn="torchvision")
            O.non_max_suppression(y, CONF, IOU, max_det=MAX_DET, use_torchvision=True)
            dt = time.perf_counter() - t0
            if it:  # first pass is the warm-up
                times.append(dt)
    t = statistics.median(times)
    return sample_images / t, t, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = 8
    reps = max(1, min(args.steps, 5))
    ips, t, threads = cpu_reference(args.workload, sample, reps)
    line = {
        "impl": "reference", "metric": "images/s @640x640 fwd+decode+NMS", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": reps, "warmup": 1, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "images_per_step": sample, "imgsz": 640, "conf": CONF, "iou": IOU,
                   "path": "oracle port of the reference's CPU path (torch CPU fp32, fused BN, torchvision nms/deform_conv2d)"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} images x {reps} passes (median), fwd+decode+NMS"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from mgdt_yolo_b200 import ops
    from mgdt_yolo_b200._lib import lib
    from mgdt_yolo_b200.engine import Engine
    from mgdt_yolo_b200.synth import raise_cls_bias, synth_state_dict
    from mgdt_yolo_b200.tasks import DetectionModel

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib()  # fail loudly now if the extension is missing

    cfg, nc, cls_bias = WORKLOADS[args.workload]
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    model = DetectionModel(cfg, nc=nc, verbose=False)
    model.load_state_dict(raise_cls_bias(synth_state_dict(model.state_dict(), seed=1), cls_bias))
    B = args.batch
    eng = Engine(model, B, 640, dtype, dev, conf=CONF, iou=IOU, max_det=MAX_DET, slots=args.slots,
                 use_graph=not args.no_graph)

    # rotating inputs: 6 x 39 MB uint8 batches on the device (> 126 MB L2) and in pinned host memory
    R = 6
    host = [make_u8(B, 100 + rank * R + i).pin_memory() for i in range(R)]
    devin = [h.to(dev) for h in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- (1) device-resident throughput: batches alternate over the engine's two slots (two CUDA streams), so
    # the tail / prologue of one batch's kernels overlaps the other's; timed from one event before the first
    # launch to one event after both streams drained.
    nslot = len(eng.slots)
    main = torch.cuda.current_stream(dev)
    for i in range(args.warmup):
        eng.step_device(devin[i % R], slot=i % nslot)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record(main)
    for sl in eng.slots:
        sl.stream.wait_event(e0)
    for i in range(args.steps):
        eng.step_device(devin[i % R], slot=i % nslot)
    for sl in eng.slots:
        main.wait_stream(sl.stream)
    e1.record(main)
    barrier()
    t_wall1 = time.time()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t_wall0, t_wall1)
    n_det = int(eng.slots[(args.steps - 1) % nslot].counts.sum().item())

    if args.launch_list:
        if rank == 0:
            print(json.dumps({"launch_list": True, "steps": args.steps, "ms_per_step": ms_dev / args.steps,
                              "launches_per_step": eng.launches_per_step}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- (2) end to end from pinned host memory (H2D + D2H inside the timed region)
    for i in range(args.warmup):
        eng.collect(eng.submit(host[i % R]))
    barrier()
    t0 = time.perf_counter()
    pending = []
    for i in range(args.steps):
        pending.append(eng.submit(host[i % R]))
        if len(pending) == len(eng.slots):
            eng.collect(pending.pop(0))
    while pending:
        eng.collect(pending.pop(0))
    torch.cuda.synchronize()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3)
    barrier()

    # ---- (3) per-launch profile of eager steps: CUDA events on the launching stream around every C-ABI call, PDL off
    # (so consecutive kernels do not overlap across the events) and the stream pre-loaded with a spin kernel so that
    # the launches are queued ahead of the GPU (the events then bracket kernel execution, not CPU launch gaps)
    prof_rows = []
    if rank == 0:
        s = eng.slots[0]
        reps = 3
        agg = {}
        lib().mgdt_set_pdl(0)
        with torch.cuda.stream(s.stream), torch.no_grad():
            for r in range(reps + 1):
                ops.PROFILE = []
                torch.cuda._sleep(6_000_000)   # ~3 ms: longer than the CPU needs to enqueue one step
                eng._head(s, devin[r % R])
                eng._body(s)
                s.stream.synchronize()
                if r:  # first pass warms the eager path
                    for k, (name, meta, a, b) in enumerate(ops.PROFILE):
                        d = agg.setdefault(k, dict(name=name, **meta, ms=0.0))
                        d["ms"] += a.elapsed_time(b) / reps
                ops.PROFILE = None
        lib().mgdt_set_pdl(1)
        prof_rows = list(agg.values())
    if world > 1:
        dist.barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    hbm, tf_burst, tf_sus, peak_src = peaks()
    total_ms = sum(r["ms"] for r in prof_rows) or 1.0
    # per-shape table (profile json) and per-KERNEL totals: the dominant kernel is the __global__ function with the
    # largest share of the step; its roofline numbers are totals over all of its launches in one step
    groups, kernels = {}, {}
    for r in prof_rows:
        g = groups.setdefault((r["name"], r["shape"]), dict(name=r["name"], shape=r["shape"], ms=0.0, n=0,
                                                            bytes=r["bytes"], flops=r["flops"]))
        g["ms"] += r["ms"]
        g["n"] += 1
        kn = r.get("kernel", r["name"].replace("mgdt_", "") + "_kernel")
        kk = kernels.setdefault(kn, dict(kernel=kn, ms=0.0, n=0, bytes=0.0, flops=0.0))
        kk["ms"] += r["ms"]; kk["n"] += 1; kk["bytes"] += r["bytes"]; kk["flops"] += r["flops"]
    top = max(kernels.values(), key=lambda g: g["ms"])
    t_launch = top["ms"] / top["n"] * 1e-3            # average launch duration
    b_launch = top["bytes"] / top["n"]                # average algorithmic bytes per launch
    f_launch = top["flops"] / top["n"]
    ai = top["flops"] / max(top["bytes"], 1)
    ridge = tf_sus * 1e12 / (hbm * 1e9)
    if ai >= ridge:
        roof = {"bound": "tensor", "achieved": f_launch / t_launch / 1e12, "peak": tf_sus, "unit": "TFLOP/s"}
    else:
        roof = {"bound": "hbm", "achieved": b_launch / t_launch / 1e9, "peak": hbm, "unit": "GB/s"}
    traffic = None
    try:  # measured DRAM bytes per launch of that kernel from the committed ncu capture (profiles/traffic_r01.json)
        with open(os.path.join(ROOT, "profiles", "traffic_r01.json")) as f:
            tj = json.load(f)
        if tj.get("kernel") == top["kernel"] and tj.get("batch") == B and tj.get("workload") == args.workload:
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        traffic = None
    roof.update(frac=roof["achieved"] / roof["peak"], traffic=traffic, kernel=top["kernel"],
                launches_per_step=top["n"], us_per_launch=t_launch * 1e6, share_of_step=top["ms"] / total_ms,
                algorithmic_bytes=b_launch, flops=f_launch, peak_source=peak_src,
                arithmetic_intensity=ai, tensor_frac_of_sustained=f_launch / t_launch / 1e12 / tf_sus,
                note="totals over all launches of the kernel in one step / their summed CUDA-event durations")

    if args.profile_json:
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_json)), exist_ok=True)
        rows = sorted(groups.values(), key=lambda g: -g["ms"])
        for g in rows:
            tl = g["ms"] / g["n"] * 1e-3
            g.update(share=g["ms"] / total_ms, gbs=g["bytes"] / tl / 1e9, tflops=g["flops"] / tl / 1e12,
                     hbm_frac=g["bytes"] / tl / 1e9 / hbm)
        with open(args.profile_json, "w") as f:
            per_kernel = sorted(kernels.values(), key=lambda g: -g["ms"])
            for g in per_kernel:
                g.update(share=g["ms"] / total_ms, us_per_launch=g["ms"] / g["n"] * 1e3,
                         gbs=g["bytes"] / (g["ms"] * 1e-3) / 1e9, hbm_frac=g["bytes"] / (g["ms"] * 1e-3) / 1e9 / hbm)
            json.dump({"workload": args.workload, "batch": B, "dtype": args.dtype, "eager_step_ms": total_ms,
                       "graph_step_ms": ms_dev / args.steps, "peaks": {"hbm_gbs": hbm, "bf16_tflops_sustained": tf_sus},
                       "per_kernel": per_kernel, "kernels": rows}, f, indent=1)

    imgs = args.steps * B * world
    line = {
        "metric": "images/s @640x640 fwd+decode+NMS", "value": imgs / (ms_dev * 1e-3), "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": args.workload, "cfg": cfg, "nc": nc, "images_per_gpu_per_step": B, "global_batch": B * world,
                   "imgsz": 640, "input": "uint8 NCHW (predictor.py:115-130), /255 on device", "conf": CONF, "iou": IOU,
                   "max_det": MAX_DET, "weights": f"random-init, synth seed 1, cls bias {cls_bias} (5% of anchors pass conf)",
                   "detections_last_step": n_det, "cuda_graph": not args.no_graph,
                   "batches_in_flight": nslot,
                   "l2": f"inputs rotate over {R} x {B * 3 * 640 * 640 / 1e6:.0f} MB batches (> 126 MB L2); "
                         "per-step activation traffic >> L2", "parallelism": f"dp{world} (batch sharded, no collective)"},
        "e2e": {"value": imgs / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": B * 3 * 640 * 640,
                "d2h_bytes_per_step": B * (MAX_DET * 6 * 4 + 4), "ms_per_step": ms_e2e / args.steps,
                "api": f"Engine.submit/collect, {nslot} batches in flight"},
        "gpu_launches": eng.launches_per_step * args.steps,
        "launches_per_step": eng.launches_per_step,
        "clocks": clocks,
        "roofline": roof,
    }
    if world == 1 and not args.no_cpu_baseline:
        ips, t, threads = cpu_reference(args.workload, 8, 3)
        line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
                                "sample": "8 images x 3 passes (median) of the same workload, fwd+decode+NMS, torch CPU fp32"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse()
    sys.exit(run_reference(a) if a.impl == "reference" else run_ours(a))
