"""non_max_suppression -- drop-in for yolo/utils/ops.py:136-266 on the batched B200 NMS kernel."""
from __future__ import annotations

import torch

from . import ops

__all__ = ("non_max_suppression", "nms_packed", "scale_boxes", "clip_boxes", "scale_boxes_params", "process_batch",
           "match_batch", "update_metrics", "IOUV")

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


def scale_boxes_params(img1_shape, img0_shape, ratio_pad=None):
    """(gain, pad_w, pad_h, h0, w0) exactly as ops.scale_boxes derives them (yolo/utils/ops.py:104-109)."""
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1), round(
            (img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1)
    else:
        gain = ratio_pad[0][0]
        pad = ratio_pad[1]
    return float(gain), float(pad[0]), float(pad[1]), float(img0_shape[0]), float(img0_shape[1])


def scale_boxes(img1_shape, boxes, img0_shape, ratio_pad=None):
    """Drop-in for ops.scale_boxes (yolo/utils/ops.py:90-117): rescales (n, >=4) xyxy boxes from the letterboxed
    image shape to the original image shape and clips them, in place, on the device (bit-exact with the reference)."""
    if boxes.shape[0] == 0:
        return boxes
    if boxes.dtype != torch.float32 or not boxes.is_contiguous() or boxes.dim() != 2:
        raise ValueError("scale_boxes: expects a contiguous fp32 (n, >=4) CUDA tensor")
    prm = torch.tensor([scale_boxes_params(img1_shape, img0_shape, ratio_pad)], dtype=torch.float32, device=boxes.device)
    ops.scale_boxes_packed(boxes.unsqueeze(0), None, prm)
    return boxes


def clip_boxes(boxes, shape):
    """ops.clip_boxes (yolo/utils/ops.py:269-285) through the same kernel (gain 1, no padding)."""
    if boxes.shape[0] == 0:
        return
    prm = torch.tensor([[1.0, 0.0, 0.0, float(shape[0]), float(shape[1])]], dtype=torch.float32, device=boxes.device)
    ops.scale_boxes_packed(boxes.unsqueeze(0), None, prm)


def IOUV(device):
    """iou vector for mAP@0.5:0.95 (yolo/v8/detect/val.py:28)."""
    return torch.linspace(0.5, 0.95, 10).to(device)


def process_batch(detections, labels, iouv=None):
    """Drop-in for DetectionValidator._process_batch (yolo/v8/detect/val.py:150-175): detections (N, 6) xyxy conf cls,
    labels (M, 5) cls xyxy, both fp32 CUDA tensors in native image space -> correct (N, niou) bool on the device."""
    iouv = IOUV(detections.device) if iouv is None else iouv.to(device=detections.device, dtype=torch.float32)
    if detections.shape[0] == 0:
        return torch.zeros((0, iouv.numel()), dtype=torch.bool, device=detections.device)
    lab = labels.to(torch.float32).contiguous().unsqueeze(0)
    return ops.match_batch(detections.to(torch.float32).contiguous().unsqueeze(0), None, lab, None, iouv.contiguous())[0]


def match_batch(dets, det_counts, labels, lab_counts, iouv=None):
    """Whole-batch form on the packed NMS output (N, max_det, 6) + counts and packed labels (N, max_lab, 5) + counts:
    one launch, no host synchronisation (the per-image loop of DetectionValidator.update_metrics, val.py:73-110)."""
    iouv = IOUV(dets.device) if iouv is None else iouv
    return ops.match_batch(dets, det_counts, labels, lab_counts, iouv.contiguous())


def update_metrics(out, counts, batch, iouv=None, single_cls=False):
    """DetectionValidator.update_metrics (yolo/v8/detect/val.py:73-110) for a whole batch on the packed NMS output.

    out (N, max_det, 6) / counts (N,) are what nms_packed returns (letterboxed-image pixels); `batch` carries the
    reference's keys: 'img' (only its shape is used), 'batch_idx' (M,), 'cls' (M, 1), 'bboxes' (M, 4) normalised xywh,
    'ori_shape' [(h0, w0)], 'ratio_pad' [((gain, gain), (pad_w, pad_h))].  Predictions and labels are mapped to native
    image space (mgdt_scale_boxes), matched at the IoU levels (mgdt_match_batch), and the per-image tuples the reference
    appends to `self.stats` -- (correct (n, niou) bool, conf, pred_cls, target_cls) -- are returned (images without
    predictions and without labels are skipped, as in the reference).  Two launches + one host read for the batch."""
    dev = out.device
    n, max_det, _ = out.shape
    height, width = int(batch['img'].shape[2]), int(batch['img'].shape[3])
    bidx = batch['batch_idx'].reshape(-1).cpu()
    cls = batch['cls'].reshape(-1, 1).float().cpu()
    bbox = batch['bboxes'].reshape(-1, 4).float().cpu()
    prm = torch.tensor([scale_boxes_params((height, width), batch['ori_shape'][si], batch['ratio_pad'][si]) for si in range(n)],
                       dtype=torch.float32, device=dev)
    predn = out.clone()
    if single_cls:
        predn[:, :, 5] = 0
    ops.scale_boxes_packed(predn, counts, prm)                       # native-space predictions (val.py:96-98)
    per_img = [(bidx == si).nonzero().reshape(-1) for si in range(n)]
    max_lab = max(1, max(int(i.numel()) for i in per_img))
    tb = torch.zeros((n, max_lab, 4), dtype=torch.float32)
    tc = torch.zeros((n, max_lab, 1), dtype=torch.float32)
    for si, idx in enumerate(per_img):
        if idx.numel():
            b = bbox[idx]
            xyxy = b.clone()                                         # ops.xywh2xyxy (yolo/utils/ops.py:372-377)
            xyxy[..., 0] = b[..., 0] - b[..., 2] / 2
            xyxy[..., 1] = b[..., 1] - b[..., 3] / 2
            xyxy[..., 2] = b[..., 0] + b[..., 2] / 2
            xyxy[..., 3] = b[..., 1] + b[..., 3] / 2
            tb[si, :idx.numel()] = xyxy * torch.tensor((width, height, width, height), dtype=torch.float32)   # val.py:103-104
            tc[si, :idx.numel()] = cls[idx]
    lab_counts = torch.tensor([int(i.numel()) for i in per_img], dtype=torch.int32, device=dev)
    tb = tb.to(dev)
    ops.scale_boxes_packed(tb, lab_counts, prm)                      # native-space labels (val.py:105-106)
    labels = torch.cat((tc.to(dev), tb), 2).contiguous()
    correct = match_batch(predn, counts, labels, lab_counts, iouv)
    cnt = counts.tolist()                                            # the single host sync
    stats = []
    for si in range(n):
        npr, nl = cnt[si], int(per_img[si].numel())
        tcls = cls[per_img[si], 0].to(dev)
        if npr == 0:
            if nl:
                stats.append((correct[si, :0], torch.zeros(0, device=dev), torch.zeros(0, device=dev), tcls))
            continue
        stats.append((correct[si, :npr], out[si, :npr, 4], predn[si, :npr, 5], tcls))
    return stats, predn
