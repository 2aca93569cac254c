"""TEST INFRASTRUCTURE -- golden-vector generator (runs ONLY in the build container).

Runs the LIVE, unmodified reference (/root/reference through oracle/ref_live.py)
on synthetic weights/inputs from mgdt_yolo_b200.synth and writes small fixtures
to tests/golden/.  The fixtures hold OUTPUTS only; weights and inputs are
regenerated from the (seed, key-name) recipe on the consumer side, so the same
file pins the oracle on CPU and the CUDA path on the GPU box.

    python oracle/make_golden.py            # regenerate everything

Fixture index (tests/golden/):
  model_<cfg>.npz    y, raw maps (B=2, 64x96) + every layer output (B=1, 64x64), un-fused eval
  modules.npz        one entry per module class on odd/ragged shapes
  nms.npz            reference non_max_suppression outputs for a parameter grid
  prepost.npz        LetterBox + BGR->RGB/CHW outputs (live cv2.resize) and ops.scale_boxes outputs
  model640.npz       the four BASELINE configs at B=2, 640x640: decoded boxes / scores + per-layer statistics
  results.npz        Boxes.xywh/xyxyn/xywhn and DetectionValidator._process_batch outputs (validation matching)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mgdt_yolo_b200.synth import synth_images, synth_predictions, synth_state_dict  # noqa: E402
from oracle import ref_live  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

from oracle.cases import (LAYER_CFGS, LETTERBOX_CASES, MODEL_CFGS, MODULE_CASES, NMS_CASES, SCALE_CASES,  # noqa: E402
                          module_inputs, synth_bgr, synth_boxes)


def gen_models():
    for cfg in MODEL_CFGS:
        m = ref_live.build_model(cfg)
        m.load_state_dict(synth_state_dict(m.state_dict(), seed=1))
        m.eval()
        blob = {}
        with torch.inference_mode():
            y, raw = m(synth_images(2, h=64, w=96, seed=0))
            blob["y"] = y.numpy()
            for i, r in enumerate(raw):
                blob[f"raw{i}"] = r.numpy()
            if cfg in LAYER_CFGS:
                x = synth_images(1, h=64, w=64, seed=5)
                ys, cur = [], x
                for layer in m.model:  # BaseModel._predict_once (nn/tasks.py:78-84), keeping every output
                    if layer.f != -1:
                        cur = ys[layer.f] if isinstance(layer.f, int) else [cur if j == -1 else ys[j] for j in layer.f]
                    cur = layer(cur)
                    ys.append(cur)
                    if isinstance(cur, torch.Tensor):
                        blob[f"layer{layer.i}"] = cur.numpy()
                blob["y1"] = cur[0].numpy()
        blob["n_params"] = np.int64(sum(p.numel() for p in m.parameters()))
        blob["keys"] = np.array(list(m.state_dict().keys()))
        blob["shapes"] = np.array([",".join(map(str, v.shape)) for v in m.state_dict().values()])
        path = os.path.join(OUT, f"model_{cfg[:-5]}.npz")
        np.savez_compressed(path, **blob)
        print(path, os.path.getsize(path) // 1024, "KiB")


def gen_model640():
    """The four BASELINE.json configs at the benchmarked size (B = 2, 640 x 640, seed 0), live reference, un-fused eval:
    decoded boxes in full; class scores in full for the nc = 2 config, else their per-anchor max / argmax and every
    16th anchor column; per-layer (mean |x|, max |x|) of every tensor-valued layer output."""
    from tests.parity import BASELINE_CFGS
    blob = {}
    for cfg, nc in BASELINE_CFGS.items():
        m = ref_live.build_model(cfg, nc=nc)
        m.load_state_dict(synth_state_dict(m.state_dict(), seed=1))
        m.eval()
        name = cfg[:-5]
        with torch.inference_mode():
            x = synth_images(2, size=640, seed=0)
            ys, cur = [], x
            for layer in m.model:
                if layer.f != -1:
                    cur = ys[layer.f] if isinstance(layer.f, int) else [cur if j == -1 else ys[j] for j in layer.f]
                cur = layer(cur)
                ys.append(cur)
            y = cur[0]
        blob[f"{name}.boxes"] = y[:, :4].numpy()
        sc = y[:, 4:]
        if nc <= 4:
            blob[f"{name}.scores"] = sc.numpy()
        else:
            blob[f"{name}.score_max"] = sc.max(1).values.numpy()
            blob[f"{name}.score_argmax"] = sc.argmax(1).to(torch.int16).numpy()
            blob[f"{name}.scores_16"] = sc[:, :, ::16].contiguous().numpy()
        blob[f"{name}.layer_stats"] = np.array([[float(t.abs().mean()), float(t.abs().max())] if isinstance(t, torch.Tensor)
                                                else [0.0, 0.0] for t in ys[:-1]], dtype=np.float64)
        print("model640", cfg, tuple(y.shape))
    path = os.path.join(OUT, "model640.npz")
    np.savez_compressed(path, **blob)
    print(path, os.path.getsize(path) // 1024, "KiB")


def gen_modules():
    ref_live.load()
    import ultralytics.nn.modules as M
    import ultralytics.nn.modules.head as H
    ns = {k: getattr(M, k) for k in dir(M)}
    ns.update(Conv_GN=H.Conv_GN, TaskDecomposition=H.TaskDecomposition)
    blob = {}
    from ultralytics.yolo.utils.torch_utils import initialize_weights
    for name, ctor, shapes, is_list in MODULE_CASES:
        torch.manual_seed(0)
        mod = eval(ctor, ns)
        initialize_weights(mod)  # BN eps = 1e-3 (torch_utils.py:254-256)
        if hasattr(mod, "stride") and isinstance(mod.stride, torch.Tensor):
            mod.stride = torch.tensor([8.0 * 2 ** i for i in range(len(shapes))])
        mod.load_state_dict(synth_state_dict(mod.state_dict(), seed=7))
        mod.eval()
        xs = module_inputs(name, shapes)
        with torch.inference_mode():
            out = mod([t.clone() for t in xs]) if is_list else mod(xs[0])
        if isinstance(out, tuple):  # heads: (y, raw)
            blob[f"{name}.y"] = out[0].numpy()
            for i, r in enumerate(out[1]):
                blob[f"{name}.raw{i}"] = r.numpy()
        else:
            blob[f"{name}.y"] = out.numpy()
        blob[f"{name}.keys"] = np.array(list(mod.state_dict().keys()))
    path = os.path.join(OUT, "modules.npz")
    np.savez_compressed(path, **blob)
    print(path, os.path.getsize(path) // 1024, "KiB")


def gen_nms():
    ref_live.load()
    from ultralytics.yolo.utils import ops
    blob = {}
    for ci, (name, nc, anchors, batch, kw) in enumerate(NMS_CASES):
        pred = synth_predictions(batch, nc, anchors, seed=20 + ci)
        out = ops.non_max_suppression(pred.clone(), max_time_img=1000.0, **kw)
        for b, t in enumerate(out):
            blob[f"{name}.{b}"] = t.numpy()
        print(name, [tuple(t.shape) for t in out])
    path = os.path.join(OUT, "nms.npz")
    np.savez_compressed(path, **blob)
    print(path, os.path.getsize(path) // 1024, "KiB")


def gen_prepost():
    """LetterBox + the BGR->RGB / HWC->CHW of BasePredictor.preprocess (predictor.py:121-125, augment.py:538-593, whose
    resize is the live cv2.resize), and ops.scale_boxes (ops.py:90-117)."""
    ref_live.load()
    from ultralytics.yolo.data.augment import LetterBox
    from ultralytics.yolo.utils import ops
    blob = {}
    for ci, (name, shape, new_shape, auto) in enumerate(LETTERBOX_CASES):
        img = synth_bgr(shape[0], shape[1], 300 + ci)
        out = LetterBox(new_shape, auto=auto, stride=32)(image=img)
        im = np.ascontiguousarray(np.stack([out])[..., ::-1].transpose((0, 3, 1, 2)))
        blob[f"lb.{name}"] = im[0]
        print("letterbox", name, shape, "->", im.shape)
    for ci, (name, s1, s0, n) in enumerate(SCALE_CASES):
        b = synth_boxes(n, s1, 400 + ci)
        out = ops.scale_boxes(s1, b.clone(), s0)
        blob[f"sb.{name}"] = out.numpy()
    path = os.path.join(OUT, "prepost.npz")
    np.savez_compressed(path, **blob)
    print(path, os.path.getsize(path) // 1024, "KiB")


def gen_results():
    """Boxes views (results.py:405-430) and DetectionValidator._process_batch (v8/detect/val.py:150-175) of the live
    reference on synthetic detections / labels."""
    import types
    ref_live.load()
    from ultralytics.yolo.engine.results import Boxes
    from ultralytics.yolo.v8.detect.val import DetectionValidator
    from oracle.cases import MATCH_CASES, synth_match
    dummy = types.SimpleNamespace(iouv=torch.linspace(0.5, 0.95, 10))
    blob = {}
    for ci, (name, nd, nl, nc, shape) in enumerate(MATCH_CASES):
        dets, labels = synth_match(nd, nl, nc, shape, 500 + ci)
        correct = DetectionValidator._process_batch(dummy, dets, labels)
        blob[f"pb.{name}"] = correct.numpy()
        b = Boxes(dets, shape)
        blob[f"xywh.{name}"], blob[f"xyxyn.{name}"], blob[f"xywhn.{name}"] = b.xywh.numpy(), b.xyxyn.numpy(), b.xywhn.numpy()
        print("match", name, tuple(correct.shape), "correct per level", correct.sum(0).tolist())
    path = os.path.join(OUT, "results.npz")
    np.savez_compressed(path, **blob)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if "--only-640" in sys.argv:
        gen_model640()
        sys.exit(0)
    if "--only-prepost" not in sys.argv and "--only-results" not in sys.argv:
        gen_models()
        gen_model640()
        gen_modules()
        gen_nms()
    if "--only-results" not in sys.argv:
        gen_prepost()
    gen_results()
