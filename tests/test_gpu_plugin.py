"""Row (b) on hardware (pytest -m gpu): the drop-in boundary exercised through the REFERENCE's own engines.

The unmodified reference travels to the GPU box as baseline/_ref (baseline/install_ref.py).  After plugin.install()
its `YOLO(yaml)` facade, `AutoBackend`, `DetectionPredictor` and `DetectionValidator` run unchanged on top of the B200
modules + libmgdt_b200.so; the same calls on the untouched reference (its own nn.Modules on cuDNN / torchvision, fp32,
TF32 off) give the expected detections and validation statistics.
"""
import pathlib

import numpy as np
import pytest
import torch

from baseline import ref_loader

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="reference copy (baseline/_ref) not present")]


def _iou(a, b):
    lt = torch.max(a[:, None, :2], b[None, :, :2])
    rb = torch.min(a[:, None, 2:4], b[None, :, 2:4])
    inter = (rb - lt).clamp(min=0).prod(-1)
    area = lambda t: (t[:, 2] - t[:, 0]) * (t[:, 3] - t[:, 1])   # noqa: E731
    return inter / (area(a)[:, None] + area(b)[None] - inter + 1e-9)


def _matched_fraction(ref, got, top=30, thr=0.9):
    """Share of the `top` most confident reference detections that have a same-class detection with IoU > thr."""
    r = ref[:top]
    if r.shape[0] == 0:
        return 1.0
    if got.shape[0] == 0:
        return 0.0
    ok = (_iou(r, got) > thr) & (r[:, 5:6] == got[None, :, 5])
    return float(ok.any(1).float().mean())


@pytest.fixture(scope="module")
def arms():
    """(reference YOLO, plugin YOLO, saved bindings): the reference model is built BEFORE install()."""
    from mgdt_yolo_b200 import plugin
    from tests import ref_pipeline as R
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref_loader.load()
    import ultralytics.yolo.engine.model  # noqa: F401  (TASK_MAP must exist before install() rebinds it)
    import ultralytics.yolo.v8.detect  # noqa: F401
    y_ref = R.make_yolo()
    saved = plugin.install()
    y_ours = R.make_yolo()
    from mgdt_yolo_b200.tasks import DetectionModel
    assert isinstance(y_ours.model, DetectionModel) and not isinstance(y_ref.model, DetectionModel)
    yield y_ref, y_ours
    plugin.uninstall(saved)


def test_predict_through_reference_predictor(arms):
    from mgdt_yolo_b200.engine import Engine
    from tests import ref_pipeline as R
    y_ref, y_ours = arms
    ims = R.synth_bgr_images(4)
    from mgdt_yolo_b200 import plugin
    saved_nms = plugin.swap_nms(False)                       # the reference arm keeps its own non_max_suppression
    try:
        d_ref = R.predict(y_ref, ims, 0, half=False)
    finally:
        plugin.swap_nms(True, saved_nms)
    assert sum(t.shape[0] for t in d_ref) > 8
    # (1) fp32 validation mode through the reference's predictor: the same detections
    d32 = R.predict(y_ours, ims, 0, half=False)
    for a, b in zip(d32, d_ref):
        # same detections up to candidates that sit within 1e-5 of the conf / IoU thresholds (fp32 summation order)
        assert abs(a.shape[0] - b.shape[0]) <= 1, f"fp32 plugin arm: {a.shape[0]} detections, reference {b.shape[0]}"
        frac = _matched_fraction(b, a, top=1000, thr=0.999)
        assert frac >= 0.97, f"fp32 plugin arm: only {frac:.3f} of the reference detections reproduced at IoU > 0.999"
        if a.shape == b.shape:
            dbox, dconf = float((a[:, :4] - b[:, :4]).abs().max()), float((a[:, 4] - b[:, 4]).abs().max())
            assert (dbox <= 0.05 and dconf <= 2e-4) or frac >= 0.97, f"fp32 plugin arm: box diff {dbox}, conf diff {dconf}"
    # (2) half=True (AutoBackend calls model.half() and feeds fp16, nn/autobackend.py:99; cast to bf16 in ops.as_act)
    # yardstick: the reference's own half mode (fp16 modules, cuDNN) against its fp32 detections -- random-init weights give
    # soft, overlapping boxes, so reduced precision moves some of them across the NMS threshold in ANY implementation
    saved_nms = plugin.swap_nms(False)
    try:
        import copy
        y_ref16 = copy.deepcopy(y_ref)
        y_ref16.predictor = None
        d_ref16 = R.predict(y_ref16, ims, 0, half=True)
    finally:
        plugin.swap_nms(True, saved_nms)
    y_half = R.make_yolo()
    d16 = R.predict(y_half, ims, 0, half=True)
    for a, b, c in zip(d16, d_ref, d_ref16):
        assert abs(a.shape[0] - b.shape[0]) <= max(3, b.shape[0] // 5), f"half arm: {a.shape[0]} detections, reference {b.shape[0]}"
        frac, frac_ref = _matched_fraction(b, a, thr=0.8), _matched_fraction(b, c, thr=0.8)
        assert frac >= min(0.9, frac_ref - 0.15), (f"half arm: {frac:.3f} of the top reference detections matched at IoU > 0.8 "
                                                   f"(the reference's own fp16 mode: {frac_ref:.3f})")
    # (3) the CUDA-graph Engine (device LetterBox, fused uint8 stem) on the same images
    # the predictor letterboxes 480 x 640 sources with auto=True, i.e. to 480 x 640 without padding; the engine's static
    # shape is set to the same rectangle (a 640 x 640 engine pads, and the global-context layers then see other statistics)
    eng = Engine(R.make_yolo().model, len(ims), (480, 640), torch.bfloat16, "cuda:0", conf=0.25, iou=0.7, slots=1)
    d_eng = [r.boxes.data.float().cpu() for r in eng.predict(ims, auto=False)]
    for a, b, c in zip(d_eng, d_ref, d_ref16):
        frac, frac_ref = _matched_fraction(b, a, thr=0.8), _matched_fraction(b, c, thr=0.8)
        assert frac >= min(0.9, frac_ref - 0.15), (f"engine arm: {frac:.3f} of the top reference detections matched at IoU > 0.8 "
                                                   f"(the reference's own fp16 mode: {frac_ref:.3f})")


def test_validator_batch_through_reference_validator(arms, tmp_path):
    from mgdt_yolo_b200 import plugin
    from tests import ref_pipeline as R
    y_ref, y_ours = arms
    nc = 2
    probe = R.synth_val_batch(4, 5, nc)
    saved_nms = plugin.swap_nms(False)
    try:
        _, _, outs = R.validate_batches(y_ref.model, [probe], "cuda:0", False, tmp_path / "probe")
        # labels = the reference's own most confident detections, so that the metrics are not trivially zero
        batch = dict(probe)
        rows, idx = [], []
        for i, det in enumerate(outs[0]):
            d = det[:8]
            xywh = torch.stack([(d[:, 0] + d[:, 2]) / 2, (d[:, 1] + d[:, 3]) / 2, d[:, 2] - d[:, 0], d[:, 3] - d[:, 1]], 1) / 640
            rows.append(torch.cat([d[:, 5:6], xywh], 1))
            idx.append(torch.full((d.shape[0],), float(i)))
        lab = torch.cat(rows)
        batch.update(cls=lab[:, :1].contiguous(), bboxes=lab[:, 1:].contiguous(), batch_idx=torch.cat(idx))
        st_ref, res_ref, _ = R.validate_batches(y_ref.model, [batch], "cuda:0", False, tmp_path / "ref")
    finally:
        plugin.swap_nms(True, saved_nms)
    st32, res32, _ = R.validate_batches(y_ours.model, [batch], "cuda:0", False, tmp_path / "ours32")
    assert res_ref["metrics/mAP50(B)"] > 0.3
    assert len(st32) == len(st_ref)
    for a, b in zip(st32, st_ref):
        assert a[0].shape == b[0].shape
        assert float((a[0] != b[0]).float().mean()) <= 0.01          # correct matrix (fp32 boxes differ by ~1e-4)
    for k in res_ref:
        assert abs(float(res32[k]) - float(res_ref[k])) <= 0.02, (k, res32[k], res_ref[k])
    # half=True: the validator's NMS regime (conf 0.001, multi_label: 12,800 candidates per image) on the bf16 path
    st16, res16, _ = R.validate_batches(R.make_yolo().model, [batch], "cuda:0", True, tmp_path / "ours16")
    assert len(st16) == len(st_ref)
    assert abs(float(res16["metrics/mAP50(B)"]) - float(res_ref["metrics/mAP50(B)"])) <= 0.1
