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
  cpu_baseline  the UNMODIFIED reference (baseline/_ref, shipped by baseline/install_ref.py) on the box's host
                cores, fp32, fused BN, all host threads, on a bounded sample (kind "reference"; the oracle port,
                kind "port", only if the copy is absent), rank 0 / N=1
  gpu_eager_baseline   the same reference model on the SAME B200 through torch eager (cuDNN / cuBLAS /
                torchvision deform_conv2d + nms kernels, fp16 channels_last) -- the library bar -- with per-op
                library timings for K1 (conv), K14 (DCNv2) and K17 (NMS) next to this repo's kernels
  sustained     the device-resident loop repeated for >= 2 s (thermal / power steady state)
  training      BASELINE configs[4]: training step of the full config, 16 images per GPU, bf16 autocast forward / backward,
                fused CUDA criterion, one flat-bucket NCCL all-reduce (N > 1), fused clip + SGD + EMA; plus the criterion
                alone against the reference's on the same GPU
  extra         batch sweep B = 1 ... 512 of the full config, the other three BASELINE configs at B = 32, NMS at the
                validator setting (conf 0.001, multi_label)

`--impl reference` times the reference's own CPU implementation alone, on this arm's config (same workload, images
per step, steps and warm-up).
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
E2E_DEPTH = int(os.environ.get("MGDT_E2E_DEPTH", "2"))   # submissions in flight per engine slot in the e2e legs (1 or 2)


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
    ap.add_argument("--no-extras", action="store_true", help="skip the sustained run, the GPU eager baseline, the sweeps")
    ap.add_argument("--sustain-seconds", type=float, default=2.0)
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
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md's clocks line).

    The driver's `steps` give a 25-40 ms timed region, shorter than one `nvidia-smi -lms` period (and nvidia-smi needs
    ~1 s to print its first row), so the samples come from NVML in this process (`pynvml`, one query pair per ~2 ms on a
    thread); `nvidia-smi -lms 100` is the fallback when NVML cannot be opened."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    PERIOD_S = 0.002

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nv, self.h, self.run = index, [], None, None, None, False
        self.max_mhz = 0.0

    def _nvml_open(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        self.nv = pynvml

    def start(self):
        try:
            self._nvml_open()
            self.run = True
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        while self.run:
            try:
                clk = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.rows.append((time.time(), clk, tuple(n for n, b in names if bits & b)))
            except Exception:
                pass
            time.sleep(self.PERIOD_S)

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, cmax = float(f[1]), float(f[2])
            except ValueError:
                continue
            self.max_mhz = max(self.max_mhz, cmax)
            rs = tuple(n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9])
                       if v.lower().startswith("active"))
            self.rows.append((time.time(), clk, rs))

    def stop(self, t0, t1):
        if self.nv is None and self.proc is None:
            return None
        if self.nv is not None:
            self.run = False
            self.t.join(timeout=1.0)
            source = "nvml"
        else:
            time.sleep(0.15)
            self.proc.terminate()
            source = "nvidia-smi"
        inside = [(clk, rs) for t, clk, rs in self.rows if t0 <= t <= t1]
        if not inside:  # region shorter than the sampling period: the samples nearest to it
            near = sorted(self.rows, key=lambda r: min(abs(r[0] - t0), abs(r[0] - t1)))[:3]
            inside = [(clk, rs) for _, clk, rs in near]
            source += " (nearest samples: none fell inside the timed region)"
        if not inside:
            return None
        reasons = set()
        for _, rs in inside:
            reasons.update(rs)
        return {"sm_mhz": statistics.median(c for c, _ in inside), "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
                "samples": len(inside), "source": source}


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


# ------------------------------------------------------------------------------------ reference arms
def synth_weights(workload, template):
    from mgdt_yolo_b200.synth import raise_cls_bias, synth_state_dict
    return raise_cls_bias(synth_state_dict(template, seed=1), WORKLOADS[workload][2])


def build_reference(workload):
    """The unmodified reference's DetectionModel (its own YAML, its own nn.Modules) with the benchmark's synthetic
    weights, eval + fuse() as AutoBackend does (nn/autobackend.py:92-101).  None if the copy did not travel."""
    from baseline import ref_loader
    if not ref_loader.available():
        return None
    ref_loader.load()
    cfg, nc, _ = WORKLOADS[workload]
    m = ref_loader.build_model(cfg, nc=nc)
    m.load_state_dict(synth_weights(workload, m.state_dict()))
    return m.eval().fuse(verbose=False)


def reference_nms(y):
    from ultralytics.yolo.utils import ops as rops
    return rops.non_max_suppression(y, CONF, IOU, max_det=MAX_DET, max_time_img=10.0)   # no wall-clock abort: every image is kept


def cpu_reference(workload, batch, steps, warmup):
    """The reference's CPU path: fp32 forward + decode + NMS of `batch` images per step with all host threads.
    -> (images/s, seconds per step (median), threads, kind)."""
    import torch
    threads = host_threads()
    torch.set_num_threads(threads)
    cfg, nc, cls_bias = WORKLOADS[workload]
    x = make_u8(batch, 0).float() / 255
    model = build_reference(workload)
    times = []
    with torch.inference_mode():
        if model is not None:
            kind = "reference"
            step = lambda: reference_nms(model(x))   # noqa: E731
        else:   # the copy is absent: the oracle restatement (pinned to the live reference by tests/test_oracle_golden.py)
            from mgdt_yolo_b200.tasks import DetectionModel
            from oracle import mgdt_oracle as O
            kind = "port"
            sd = O.fold_bn(synth_weights(workload, DetectionModel(cfg, nc=nc, verbose=False).state_dict()))

            def step():
                y, _, _ = O.forward(cfg, sd, x, nc=nc, dcn="torchvision")
                return O.non_max_suppression(y, CONF, IOU, max_det=MAX_DET, use_torchvision=True)
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            step()
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    t = statistics.median(times)
    return batch / t, t, threads, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ips, t, threads, kind = cpu_reference(args.workload, args.batch, args.steps, args.warmup)
    what = ("the unmodified reference (baseline/_ref): DetectionModel.fuse() fp32 + ops.non_max_suppression on the host cores"
            if kind == "reference" else "oracle port of the reference's CPU path (torch CPU fp32, fused BN, torchvision nms/deform_conv2d)")
    line = {
        "impl": "reference", "metric": "images/s @640x640 fwd+decode+NMS", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "images_per_gpu_per_step": args.batch, "imgsz": 640, "conf": CONF, "iou": IOU,
                   "max_det": MAX_DET, "path": what},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": kind,
                         "sample": f"{args.batch} images x {args.steps} steps (median step, {args.warmup} warm-up), fwd+decode+NMS"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def gpu_eager_reference(workload, batch, dev, steps=10, warmup=3):
    """The library bar: the reference model itself on this B200 through torch eager -- cuDNN convolutions, cuBLAS
    linears, torchvision's deform_conv2d and nms CUDA kernels -- fp16 (the reference's half mode, autobackend.py:99)
    in channels_last, forward + decode + its own non_max_suppression.  Also per-op library timings."""
    import torch
    import torch.nn.functional as F
    import torchvision
    model = build_reference(workload)
    if model is None:
        return None
    out = {"dtype": "fp16 (model NCHW as the reference runs it: its heads .view() the maps; per-op convs channels_last)", "path": "reference nn.Modules, torch eager (cuDNN / cuBLAS / torchvision CUDA ops)",
           "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
    model = model.to(dev).half()
    xs = [(make_u8(batch, 200 + i).to(dev).half() / 255).contiguous() for i in range(3)]

    def timed(fn, n=steps, w=warmup):
        for _ in range(w):
            fn(0)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n

    with torch.inference_mode():
        ms_fwd = timed(lambda i: model(xs[i % 3]))
        y = model(xs[0])[0]
        ms_nms = timed(lambda i: reference_nms(y))
        ms_all = timed(lambda i: reference_nms(model(xs[i % 3])))
        out.update(value=batch / (ms_all * 1e-3), unit="images/s", images_per_step=batch, ms_per_step=ms_all,
                   ms_forward_decode=ms_fwd, ms_nms=ms_nms)
        # per-op library timings on the shapes of the full config (B = 32): K1 cuDNN conv (+ SiLU), K14 torchvision
        # deform_conv2d, K17 the reference's non_max_suppression (torchvision.ops.nms inside a per-image python loop)
        ops_ = []
        for cin, cout, k, hw in ((32, 32, 3, 80), (64, 256, 1, 80), (96, 384, 1, 40), (16, 32, 3, 160)):
            s2 = 2 if (cin, cout) == (16, 32) else 1
            xin = torch.randn(batch, cin, hw * s2, hw * s2, device=dev, dtype=torch.half).contiguous(memory_format=torch.channels_last)
            w = torch.randn(cout, cin, k, k, device=dev, dtype=torch.half).contiguous(memory_format=torch.channels_last)
            b = torch.randn(cout, device=dev, dtype=torch.half)
            ms = timed(lambda i: F.silu(F.conv2d(xin, w, b, s2, k // 2)), 20)
            ops_.append({"op": "K1 cuDNN conv2d+bias, SiLU", "shape": f"{cin}->{cout} k{k}s{s2} {batch}x{hw * s2}x{hw * s2}", "lib_us": ms * 1e3})
        xin = torch.randn(batch, 32, 80, 80, device=dev, dtype=torch.half)
        off = torch.randn(batch, 18, 80, 80, device=dev, dtype=torch.half)
        msk = torch.rand(batch, 9, 80, 80, device=dev, dtype=torch.half)
        w = torch.randn(32, 32, 3, 3, device=dev, dtype=torch.half)
        try:
            ms = timed(lambda i: torchvision.ops.deform_conv2d(xin, off, w, None, 1, 1, 1, msk), 10)
            ops_.append({"op": "K14 torchvision.ops.deform_conv2d", "shape": f"dcn 32->32 {batch}x80x80", "lib_us": ms * 1e3})
        except Exception as e:   # noqa: BLE001
            ops_.append({"op": "K14 torchvision.ops.deform_conv2d", "shape": f"dcn 32->32 {batch}x80x80", "error": str(e)[:80]})
        ops_.append({"op": "K17 reference non_max_suppression (torchvision.ops.nms)", "shape": f"nms N{batch} nc{y.shape[1] - 4} A{y.shape[2]}",
                     "lib_us": ms_nms * 1e3})
        out["library_ops"] = ops_
    del model, xs
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------ GPU arm
def build_engine(workload, batch, dtype, dev, slots, use_graph=True, **nms):
    from mgdt_yolo_b200.engine import Engine
    from mgdt_yolo_b200.tasks import DetectionModel
    cfg, nc, _ = WORKLOADS[workload]
    model = DetectionModel(cfg, nc=nc, verbose=False)
    model.load_state_dict(synth_weights(workload, model.state_dict()))
    kw = dict(conf=CONF, iou=IOU, max_det=MAX_DET)
    kw.update(nms)
    if os.environ.get("MGDT_COPY_SPLIT"):
        kw["copy_split"] = int(os.environ["MGDT_COPY_SPLIT"])   # A/B: concurrent H2D chunk copies per batch
    return Engine(model, batch, 640, dtype, dev, slots=slots, use_graph=use_graph, **kw)


def device_loop(eng, devin, steps, warmup, dev):
    """`steps` device-resident batches alternating over the engine's slots -> (ms total, ms e2e is measured elsewhere)."""
    import torch
    nslot, R = len(eng.slots), len(devin)
    main = torch.cuda.current_stream(dev)
    for i in range(warmup):
        eng.step_device(devin[i % R], slot=i % nslot)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for i in range(steps):
        eng.step_device(devin[i % R], slot=i % nslot)
    for sl in eng.slots:
        main.wait_stream(sl.stream)
    e1.record(main)
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1)


def e2e_loop(eng, host, steps, warmup):
    import torch
    R = len(host)
    for i in range(warmup):
        eng.collect(eng.submit(host[i % R]))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pending = []
    for i in range(steps):
        pending.append(eng.submit(host[i % R]))
        if len(pending) == E2E_DEPTH * len(eng.slots):   # two submissions in flight per slot (the next input is copied while the current batch computes)
            eng.collect(pending.pop(0))
    while pending:
        eng.collect(pending.pop(0))
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3


def extras(args, dev, dtype):
    """Bounded side measurements of the default N = 1 run (each engine is built, timed and freed)."""
    import torch
    from mgdt_yolo_b200 import ops
    out = {}
    # (a) BASELINE.json configs[3]: batch sweep of the full config
    sweep = []
    for B in (1, 2, 4, 8, 16, 64, 128, 256, 512):
        try:
            slots = 3 if B <= 128 else 2
            eng = build_engine(args.workload, B, dtype, dev, slots)
            R = max(2, min(6, 160 // max(1, B * 3 * 640 * 640 // 1000000) + 1))
            host = [make_u8(B, 300 + i).pin_memory() for i in range(R)]
            devin = [h.to(dev) for h in host]
            steps = max(6, min(60, 2048 // B))
            ms = device_loop(eng, devin, steps, 3, dev)
            ms2 = e2e_loop(eng, host, steps, 3)
            sweep.append({"batch": B, "value": steps * B / (ms * 1e-3), "e2e": steps * B / (ms2 * 1e-3), "ms_per_step": ms / steps,
                          "steps": steps, "launches_per_step": eng.launches_per_step})
            del eng, host, devin
            torch.cuda.empty_cache()
        except Exception as e:   # noqa: BLE001
            sweep.append({"batch": B, "error": str(e)[:120]})
    out["batch_sweep"] = {"workload": args.workload, "unit": "images/s", "rows": sweep}
    # (b) the other BASELINE configs at B = 32 (configs[1], configs[2], configs[0]'s graph on the GPU)
    others = []
    for wl in ("mspa_c2f_yolov8n", "mspa_c2f_gd_yolov8n", "yolov8n"):
        try:
            eng = build_engine(wl, 32, dtype, dev, 3)
            host = [make_u8(32, 400 + i).pin_memory() for i in range(6)]
            devin = [h.to(dev) for h in host]
            ms = device_loop(eng, devin, 20, 3, dev)
            ms2 = e2e_loop(eng, host, 20, 3)
            others.append({"workload": wl, "batch": 32, "value": 20 * 32 / (ms * 1e-3), "e2e": 20 * 32 / (ms2 * 1e-3),
                           "ms_per_step": ms / 20, "launches_per_step": eng.launches_per_step})
            del eng, host, devin
            torch.cuda.empty_cache()
        except Exception as e:   # noqa: BLE001
            others.append({"workload": wl, "error": str(e)[:120]})
    out["other_configs"] = others
    # (c) NMS alone on synthetic predictions: the predictor setting and the validator setting
    from mgdt_yolo_b200.synth import synth_predictions
    nms_rows = []
    for name, kw in (("predict conf0.25 single-label", dict(conf_thres=0.25, iou_thres=0.7)),
                     ("validator conf0.001 multi_label", dict(conf_thres=0.001, iou_thres=0.7, multi_label=True))):
        pred = synth_predictions(32, 2, 6400, seed=2).to(dev)
        o, c = ops.nms_packed(pred, max_det=MAX_DET, **kw)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.nms_packed(pred, max_det=MAX_DET, out=o, counts=c, **kw)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / 10
        thr = kw["conf_thres"]
        sc = pred[:, 4:]
        ncand = int(((sc > thr).sum() if kw.get("multi_label") else (sc.amax(1) > thr).sum()).item())
        per_img = ncand / 32
        nms_rows.append({"setting": name, "batch": 32, "candidates_per_image": per_img, "ms": ms,
                         "candidates_per_s": ncand / (ms * 1e-3), "iou_pairs_per_s": 32 * per_img * per_img / 2 / (ms * 1e-3),
                         "kept": int(c.sum().item())})
    out["nms"] = nms_rows
    return out


def training_leg(dev, world, rank, barrier, max_over_ranks, steps=5, warmup=2, batch=16):
    """BASELINE.json configs[4] / SURVEY 8(d) #5: training step of the full MGDT config, 16 images per GPU at 640 x 640,
    bf16 autocast forward / backward (train-mode modules: torch operators), the fused CUDA criterion, ONE flat-bucket
    gradient all-reduce over NCCL, fused clip + SGD + EMA.  Every rank takes part; timed on the device, max over ranks."""
    import torch
    from mgdt_yolo_b200.tasks import DetectionModel
    from mgdt_yolo_b200.train import FlatBucket, synth_targets, train_step, v8DetectionLoss
    wl = "mspa_c2f_gd_tood_yolov8n"
    cfg, nc, _ = WORKLOADS[wl]
    m = DetectionModel(cfg, nc=nc, verbose=False)
    m.load_state_dict(synth_weights(wl, m.state_dict()))
    m = m.to(dev)
    fb = FlatBucket(m, lr=0.01)
    b = synth_targets(batch, 20, nc, seed=rank, device=dev)
    b["img"] = make_u8(batch, 500 + rank).to(dev)
    for _ in range(warmup):
        train_step(m, fb, b)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss, items = train_step(m, fb, b)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    out = {"workload": wl, "images_per_gpu_per_step": batch, "imgsz": 640, "n_gpus": world, "steps": steps, "warmup": warmup,
           "value": steps * batch * world / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms / steps, "dtype": "bf16 autocast",
           "allreduce": {"backend": "nccl" if world > 1 else None, "calls_per_step": 1 if world > 1 else 0, "bytes": fb.n * 4},
           "params": fb.n, "loss_last": float(loss), "forward_backward": "torch operators (train_forward.py)",
           "native": ["mgdt_v8_loss", "mgdt_sumsq", "mgdt_sgd_step", "mgdt_ema_update"]}
    if rank == 0:
        # the criterion alone at this step's shapes: the fused CUDA call vs the reference's own v8DetectionLoss on this GPU
        from oracle.train_cases import model_stub
        f = torch.randn(batch, nc + 64, 80, 80, device=dev)
        f[:, :64] *= 1.5

        def timed(crit, n=10):
            x = f.clone().requires_grad_(True)
            for _ in range(2):
                crit([x], b)[0].backward()
            torch.cuda.synchronize(dev)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(n):
                crit([x], b)[0].backward()
            a1.record()
            torch.cuda.synchronize(dev)
            return a0.elapsed_time(a1) / n
        stub = model_stub(nc, 16, [(80, 80, 8.0)], dev)
        out["criterion_ms"] = {"ours_fwd_bwd": timed(v8DetectionLoss(stub))}
        try:
            from baseline import ref_loader
            # N = 1 only: importing the reference under an initialised process group runs a rank-0-only dist.barrier()
            # (yolo/utils/__init__.py SettingsManager -> torch_distributed_zero_first), which would unpair the ranks
            if world == 1 and ref_loader.available():
                ref_loader.load()
                from ultralytics.yolo.utils import loss as L
                out["criterion_ms"]["reference_fwd_bwd"] = timed(L.v8DetectionLoss(stub))
        except Exception as e:   # noqa: BLE001
            out["criterion_ms"]["reference_error"] = str(e)[:120]
    del m, fb, b
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mgdt_yolo_b200 import ops
    from mgdt_yolo_b200._lib import lib
    from mgdt_yolo_b200.parallel import ShardedEngine, bind_rank_to_cores

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        # a rank that drops out must not park the others for NCCL's default 10 minutes
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=int(os.environ.get("MGDT_BENCH_NCCL_TIMEOUT", "300"))))
    lib()  # fail loudly now if the extension is missing

    cfg, nc, cls_bias = WORKLOADS[args.workload]
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    B = args.batch
    cores = bind_rank_to_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", world))) if world > 1 else None
    eng = build_engine(args.workload, B, dtype, dev, args.slots, use_graph=not args.no_graph)

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

    # ---- (1b) sustained: the same loop repeated for >= --sustain-seconds (power / thermal steady state; the driver's
    # `steps` give a 25-40 ms timed region, shorter than the board's power controller reacts)
    sustained = None
    if not args.no_extras:
        n_sus = max(args.steps, int(args.sustain_seconds * 1e3 / max(ms_dev / args.steps, 1e-3)) + 1)
        sampler2 = ClockSampler(local)
        sampler2.start()
        time.sleep(0.25)
        barrier()
        tw0 = time.time()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(main)
        for i in range(n_sus):
            eng.step_device(devin[i % R], slot=i % nslot)
        for sl in eng.slots:
            main.wait_stream(sl.stream)
        s1.record(main)
        barrier()
        tw1 = time.time()
        ms_sus = max_over_ranks(s0.elapsed_time(s1))
        sustained = {"value": n_sus * B * world / (ms_sus * 1e-3), "unit": "images/s", "steps": n_sus, "seconds": ms_sus * 1e-3,
                     "ms_per_step": ms_sus / n_sus, "clocks": sampler2.stop(tw0, tw1)}

    # ---- (2) end to end from pinned host memory (H2D + D2H inside the timed region).  N > 1: through the sharding API
    # (parallel.ShardedEngine: global batch -> this rank's slice -> packed results gathered on rank 0 in image order)
    runner = ShardedEngine(eng, B * world) if world > 1 else eng
    for i in range(args.warmup):
        runner.collect(runner.submit(host[i % R]))
    barrier()
    t0 = time.perf_counter()
    pending = []
    n_out = 0
    for i in range(args.steps):
        pending.append(runner.submit(host[i % R]))
        if len(pending) == (E2E_DEPTH if world == 1 else 1) * len(eng.slots):
            r_ = runner.collect(pending.pop(0))
            n_out += len(r_) if r_ is not None else 0
    while pending:
        r_ = runner.collect(pending.pop(0))
        n_out += len(r_) if r_ is not None else 0
    torch.cuda.synchronize()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3)
    barrier()
    if rank == 0 and n_out != args.steps * B * world:
        raise SystemExit(f"bench.py: the e2e leg returned {n_out} per-image results, expected {args.steps * B * world}")

    # ---- (2b) BASELINE configs[4]: the training step (all ranks: its gradient all-reduce is the one collective of the repo)
    training = None
    if not args.no_extras:
        ok = torch.ones(1, device=dev)
        try:
            training = training_leg(dev, world, rank, barrier, max_over_ranks)
        except Exception as e:   # noqa: BLE001
            if world > 1:
                raise            # a rank that drops out of the collective would hang the others
            training = {"error": str(e)[:200]}
        del ok

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
    try:  # measured DRAM bytes per launch of that kernel from the committed ncu capture (profiles/traffic_rNN.json)
        for fn in ("traffic_r02.json", "traffic_r01.json"):
            path = os.path.join(ROOT, "profiles", fn)
            if not os.path.isfile(path):
                continue
            with open(path) as f:
                tj = json.load(f)
            tj = tj.get("kernels", {}).get(top["kernel"], tj)
            if tj.get("kernel") == top["kernel"] and tj.get("workload") == args.workload and tj.get("batch") == B:
                traffic = tj["dram_bytes_per_launch"]
                break
    except Exception:   # noqa: BLE001
        traffic = None
    roof.update(frac=roof["achieved"] / roof["peak"], traffic=traffic, kernel=top["kernel"],
                launches_per_step=top["n"], us_per_launch=t_launch * 1e6, share_of_step=top["ms"] / total_ms,
                algorithmic_bytes=b_launch, flops=f_launch, peak_source=peak_src,
                arithmetic_intensity=ai, tensor_frac_of_sustained=f_launch / t_launch / 1e12 / tf_sus,
                note="totals over all launches of the kernel in one step / their summed CUDA-event durations",
                # the other kernels of the step, same definition (share of the serialised step, algorithmic GB/s / measured peak)
                other_kernels=[{"kernel": g["kernel"], "launches_per_step": g["n"], "share_of_step": g["ms"] / total_ms,
                                "us_per_launch": g["ms"] / g["n"] * 1e3, "hbm_frac": g["bytes"] / (g["ms"] * 1e-3) / 1e9 / hbm}
                               for g in sorted(kernels.values(), key=lambda g: -g["ms"]) if g is not top][:8])

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
                "api": (f"ShardedEngine.submit/collect over {world} ranks (results gathered on rank 0 in image order), " if world > 1 else "")
                       + f"Engine.submit/collect, {nslot} slots x {E2E_DEPTH if world == 1 else 1} submissions in flight", "host_cores_bound": cores},
        "gpu_launches": eng.launches_per_step * args.steps,
        "launches_per_step": eng.launches_per_step,
        "clocks": clocks if clocks is not None else (dict(sustained["clocks"], source="sustained leg") if sustained and sustained.get("clocks") else None),
        "roofline": roof,
    }
    if world == 1:
        # the platform's pinned host -> device rate with as many copies in flight as the e2e leg has batches: the
        # ceiling of an end-to-end images/s that moves 1.2 MB per image over PCIe
        try:
            cs = [torch.cuda.Stream(device=dev) for _ in range(nslot)]
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            reps = 4 * nslot
            for i in range(reps):
                with torch.cuda.stream(cs[i % nslot]):
                    devin[i % R].copy_(host[i % R], non_blocking=True)
            torch.cuda.synchronize(dev)
            gbs = reps * B * 3 * 640 * 640 / (time.perf_counter() - t0) / 1e9
            line["e2e"].update(h2d_gbs_measured=gbs, h2d_ceiling_images_per_s=gbs * 1e9 / (3 * 640 * 640),
                               frac_of_h2d_ceiling=line["e2e"]["value"] / (gbs * 1e9 / (3 * 640 * 640)))
        except Exception:   # noqa: BLE001
            pass
    if sustained is not None:
        line["sustained"] = sustained
    if training is not None:
        line["training"] = training
    if world == 1 and not args.no_extras:
        try:
            ge = gpu_eager_reference(args.workload, B, dev)
            if ge is not None:
                mine = {(r["name"], r["shape"]): r["ms"] / r["n"] * 1e3 for r in groups.values()}
                for o in ge.get("library_ops", []):
                    for (nm, shp), us in mine.items():
                        if shp == o["shape"]:
                            o["ours_us"] = us
                ge["speedup_value_over_eager"] = line["value"] / ge["value"]
                line["gpu_eager_baseline"] = ge
        except Exception as e:   # noqa: BLE001
            line["gpu_eager_baseline"] = {"error": str(e)[:200]}
        try:
            line["extra"] = extras(args, dev, dtype)
        except Exception as e:   # noqa: BLE001
            line["extra"] = {"error": str(e)[:200]}
    if world == 1 and not args.no_cpu_baseline:
        ips, t, threads, kind = cpu_reference(args.workload, 8, 3, 1)
        line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": threads, "kind": kind,
                                "sample": "8 images x 3 steps (median, 1 warm-up) of the same workload, fwd+decode+NMS, torch CPU fp32"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse()
    sys.exit(run_reference(a) if a.impl == "reference" else run_ours(a))
