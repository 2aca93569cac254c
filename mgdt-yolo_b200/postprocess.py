"""non_max_suppression -- drop-in for yolo/utils/ops.py:136-266 on the batched B200 NMS kernel."""
from __future__ import annotations

import torch

from . import ops

__all__ = ("non_max_suppression", "nms_packed")

nms_packed = ops.nms_packed


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                        labels=(), max_det=300, nc=0, max_time_img=0.05, max_nms=30000, max_wh=7680):
    """Same signature and return type as the reference: a list of (n_i, 6) tensors
    (x1, y1, x2, y2, confidence, class) per image.

    Differences, all documented in DESIGN.md: the whole batch runs in four launches with ONE
    device->host read (the per-image counts) instead of a Python loop with a sync per image; there
    is no wall-clock abort (`max_time_img` is accepted and ignored, so no image is ever silently
    dropped, ops.py:262-264); score ties are ordered by candidate index (the reference's argsort
    is unstable); mask channels (nm > 0) and apriori `labels` belong to the segmentation /
    autolabel paths and are not on this hot path.
    """
    assert 0 <= conf_thres <= 1, f'Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0'
    assert 0 <= iou_thres <= 1, f'Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0'
    if isinstance(prediction, (list, tuple)):  # (inference_out, loss_out) (ops.py:183-184)
        prediction = prediction[0]
    if labels:
        raise NotImplementedError("non_max_suppression: apriori `labels` (autolabelling) are not on the B200 hot path")
    ch = prediction.shape[1]
    nc = nc or (ch - 4)
    if ch - nc - 4 != 0:
        raise NotImplementedError("non_max_suppression: mask channels (nm > 0) belong to the segmentation task")
    bs = prediction.shape[0]
    if bs == 0:
        return []
    out, counts = ops.nms_packed(prediction, conf_thres, iou_thres, multi_label=multi_label, agnostic=agnostic,
                                 max_det=max_det, max_nms=max_nms, max_wh=float(max_wh), classes=classes)
    cnt = counts.tolist()  # the single host sync of the whole call
    return [out[i, :cnt[i]] for i in range(bs)]
