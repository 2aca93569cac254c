"""TEST INFRASTRUCTURE: drive the REFERENCE's own engines (YOLO facade -> DetectionPredictor, AutoBackend,
DetectionValidator) on a given model -- the unmodified reference model or, after plugin.install(), the B200 modules
built by the reference's own parse_model.  The code below only calls the reference's public API
(yolo/engine/model.py:222-248, yolo/engine/predictor.py:210-248, yolo/v8/detect/val.py:30-150)."""
from __future__ import annotations

import numpy as np
import torch

from baseline import ref_loader
from mgdt_yolo_b200.synth import raise_cls_bias, synth_state_dict

CFG = "mspa_c2f_gd_tood_yolov8n.yaml"


def yaml_path(cfg=CFG):
    import os
    return os.path.join(ref_loader.REFERENCE_DIR, "models", "v8", cfg)


def synth_bgr_images(n, h=480, w=640, seed=11):
    rng = np.random.default_rng(seed)
    return [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(n)]


def make_yolo(cfg=CFG, cls_bias=-1.238, seed=1):
    """YOLO(<reference yaml>) with the synthetic weights of the benchmark (whatever DetectionModel class TASK_MAP holds)."""
    ref_loader.load()
    from ultralytics import YOLO
    y = YOLO(yaml_path(cfg))
    y.model.load_state_dict(raise_cls_bias(synth_state_dict(y.model.state_dict(), seed=seed), cls_bias))
    return y


def predict(yolo, images, device, half, conf=0.25, iou=0.7, imgsz=640):
    """-> list of (n_i, 6) CPU fp32 tensors (xyxy in ORIGINAL image pixels, conf, cls) from yolo.predict."""
    res = yolo.predict(images, device=device, half=half, conf=conf, iou=iou, imgsz=imgsz, verbose=False, save=False)
    return [r.boxes.data.detach().float().cpu() for r in res]


def synth_val_batch(n, labels_per_img, nc, seed=3, size=640):
    """A validation batch as the reference's dataloader collates it (yolo/data/dataset.py collate_fn): uint8 images,
    normalised xywh labels, batch_idx, ori_shape, ratio_pad."""
    g = torch.Generator().manual_seed(seed)
    img = torch.randint(0, 256, (n, 3, size, size), dtype=torch.uint8, generator=g)
    nl = n * labels_per_img
    cxy = torch.rand(nl, 2, generator=g) * 0.8 + 0.1
    wh = torch.rand(nl, 2, generator=g) * 0.25 + 0.03
    return {"img": img, "cls": torch.randint(0, nc, (nl, 1), generator=g).float(), "bboxes": torch.cat([cxy, wh], 1),
            "batch_idx": torch.arange(n).repeat_interleave(labels_per_img).float(),
            "ori_shape": [(size, size)] * n, "ratio_pad": [((1.0, 1.0), (0.0, 0.0))] * n,
            "im_file": [f"synthetic_{i}.jpg" for i in range(n)]}


def validate_batches(model, batches, device, half, save_dir, conf=0.001, iou=0.7):
    """DetectionValidator on in-memory batches: AutoBackend(model) -> preprocess -> model -> postprocess (NMS at the
    validator setting: conf 0.001, multi_label) -> update_metrics -> get_stats.  Returns (stats tuples, results_dict,
    per-batch NMS outputs)."""
    ref_loader.load()
    from ultralytics.nn.autobackend import AutoBackend
    from ultralytics.yolo.cfg import get_cfg
    from ultralytics.yolo.utils import DEFAULT_CFG
    from ultralytics.yolo.v8.detect.val import DetectionValidator
    args = get_cfg(DEFAULT_CFG, dict(mode="val", conf=conf, iou=iou, half=half, plots=False, save_json=False, verbose=False))
    v = DetectionValidator(save_dir=save_dir, args=args)
    v.device = torch.device(device)
    v.training = False
    v.data = {"val": "", "names": model.names}
    v.model = AutoBackend(model, device=v.device, fp16=half, fuse=True, verbose=False)
    v.model.eval()
    v.init_metrics(v.model)
    outs = []
    with torch.inference_mode():
        for b in batches:
            b = v.preprocess({k: (t.clone() if isinstance(t, torch.Tensor) else t) for k, t in b.items()})
            preds = v.postprocess(v.model(b["img"]))
            v.update_metrics(preds, b)
            outs.append([p.detach().float().cpu() for p in preds])
    stats = [tuple(t.detach().cpu() for t in s) for s in v.stats]
    return stats, v.get_stats(), outs
