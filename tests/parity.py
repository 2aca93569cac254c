"""Shared parity helpers: run the B200 path and measure its distance to golden vectors / the oracle.

Error metric for floating-point maps: max|a-b| / max|b|  (relative to the tensor's magnitude, as
SURVEY.md §9.13 requires) and the relative L2 error.  Tolerances (BASELINE.json north_star):
fp32 validation mode 1e-4, bf16 1e-2.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from mgdt_yolo_b200.synth import synth_images, synth_predictions, synth_state_dict
from oracle.cases import MODULE_CASES, NMS_CASES, module_inputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = {torch.float32: 1e-4, torch.bfloat16: 1e-2}


def errs(a, b):
    a = torch.as_tensor(a).detach().float().cpu()
    b = torch.as_tensor(b).detach().float().cpu()
    if a.shape != b.shape:
        return float("inf"), float("inf")
    if b.numel() == 0:
        return 0.0, 0.0
    d = (a - b)
    return float(d.abs().max()) / max(float(b.abs().max()), 1e-6), float(d.norm()) / max(float(b.norm()), 1e-6)


def build_module(ctor):
    import mgdt_yolo_b200.modules as M
    return eval(ctor, {k: getattr(M, k) for k in M.__all__})


def run_module_case(case, dtype, device="cuda"):
    """-> dict name -> (max_rel, l2_rel) against tests/golden/modules.npz."""
    name, ctor, shapes, is_list = case
    g = np.load(os.path.join(GOLDEN, "modules.npz"))
    mod = build_module(ctor)
    mod.load_state_dict(synth_state_dict(mod.state_dict(), seed=7))
    if hasattr(mod, "stride") and isinstance(mod.stride, torch.Tensor):
        mod.stride = torch.tensor([8.0 * 2 ** i for i in range(len(shapes))])
    mod = mod.to(device).eval()
    xs = [t.to(device=device, dtype=dtype) for t in module_inputs(name, shapes)]
    with torch.no_grad():
        out = mod(list(xs)) if is_list else mod(xs[0])
    res = {}
    if isinstance(out, tuple):
        res["y"] = errs(out[0], g[f"{name}.y"])
        for i, r in enumerate(out[1]):
            res[f"raw{i}"] = errs(r, g[f"{name}.raw{i}"])
    else:
        res["y"] = errs(out, g[f"{name}.y"])
    return res


def build_model(cfg, device="cuda", seed=1, nc=None, cls_bias=None):
    from mgdt_yolo_b200.synth import raise_cls_bias
    from mgdt_yolo_b200.tasks import DetectionModel
    m = DetectionModel(cfg, nc=nc, verbose=False)
    sd = synth_state_dict(m.state_dict(), seed=seed)
    if cls_bias is not None:
        sd = raise_cls_bias(sd, cls_bias)
    m.load_state_dict(sd)
    return m.to(device).eval(), sd


def run_model_case(cfg, dtype, device="cuda", layers=False):
    g = np.load(os.path.join(GOLDEN, f"model_{cfg[:-5]}.npz"))
    m, _ = build_model(cfg, device)
    res = {}
    with torch.no_grad():
        y, raw = m(synth_images(2, h=64, w=96, seed=0).to(device=device, dtype=dtype))
        res["y"] = errs(y, g["y"])
        for i, r in enumerate(raw):
            res[f"raw{i}"] = errs(r, g[f"raw{i}"])
        if layers and "y1" in g:
            x = synth_images(1, h=64, w=64, seed=5).to(device=device, dtype=dtype)
            ys, cur = [], x
            for layer in m.model:
                if layer.f != -1:
                    cur = ys[layer.f] if isinstance(layer.f, int) else [cur if j == -1 else ys[j] for j in layer.f]
                cur = layer(cur)
                ys.append(cur)
                if isinstance(cur, torch.Tensor) and f"layer{layer.i}" in g:
                    res[f"layer{layer.i}:{type(layer).__name__}"] = errs(cur, g[f"layer{layer.i}"])
            res["y1"] = errs(cur[0], g["y1"])
    return res


def run_nms_case(ci, device="cuda"):
    """-> list of (equal, n_out, n_ref) per image, against the reference's golden NMS output."""
    from mgdt_yolo_b200.postprocess import non_max_suppression
    name, nc, anchors, batch, kw = NMS_CASES[ci]
    g = np.load(os.path.join(GOLDEN, "nms.npz"))
    pred = synth_predictions(batch, nc, anchors, seed=20 + ci).to(device)
    out = non_max_suppression(pred, **kw)
    res = []
    for b, t in enumerate(out):
        ref = torch.as_tensor(g[f"{name}.{b}"])
        t = t.cpu()
        res.append((t.shape == ref.shape and bool(torch.equal(t, ref)), t.shape[0], ref.shape[0]))
    return res
