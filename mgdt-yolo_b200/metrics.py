"""Tail of row f2: the metric aggregation after the matching (DetectionValidator.get_stats, yolo/v8/detect/val.py:124-150).

`postprocess.update_metrics` leaves the validator's per-image `stats` tuples (correct matrix, confidence, predicted
class, target class) on the device; this module reduces them to the reference's numbers.  It runs ONCE per validation
pass on a few thousand rows, on the host in numpy like the reference does (yolo/utils/metrics.py) -- not a kernel.

  compute_ap        yolo/utils/metrics.py:377-406   101-point interpolated AP of one precision / recall curve
  ap_per_class      yolo/utils/metrics.py:409-504   (without the plots)
  Metric, DetMetrics  :507-717                       p / r / f1 / all_ap containers, `results_dict`, `fitness`
  ConfusionMatrix   :177-290 (detect task)           process_batch on (x1, y1, x2, y2, conf, cls) / (cls, x1, y1, x2, y2)
  get_stats         val.py:124-131                   list of per-image tuples -> results dict
"""
from __future__ import annotations

import numpy as np
import torch

__all__ = ("compute_ap", "ap_per_class", "Metric", "DetMetrics", "ConfusionMatrix", "get_stats")


def _box_filter(y, f=0.05):
    """smooth() of metrics.py:319-324: box filter over a fraction f of the curve, edge-padded."""
    nf = round(len(y) * f * 2) // 2 + 1
    pad = np.ones(nf // 2)
    return np.convolve(np.concatenate((pad * y[0], y, pad * y[-1]), 0), np.ones(nf) / nf, mode="valid")


def compute_ap(recall, precision):
    """-> (ap, precision envelope, recall with sentinels)."""
    mrec = np.concatenate(([0.0], recall, [1.0]))
    mpre = np.concatenate(([1.0], precision, [0.0]))
    mpre = np.flip(np.maximum.accumulate(np.flip(mpre)))
    x = np.linspace(0, 1, 101)
    y = np.interp(x, mrec, mpre)
    ap = float(np.sum((y[1:] + y[:-1]) * np.diff(x)) / 2)          # trapezoid rule (np.trapz)
    return ap, mpre, mrec


def ap_per_class(tp, conf, pred_cls, target_cls, eps=1e-16, **_unused):
    """-> tp, fp, p, r, f1, ap (classes with labels x IoU levels), unique_classes -- at the confidence that maximises
    the smoothed mean F1.  `plot`, `names`, ... of the reference's signature are accepted and ignored."""
    order = np.argsort(-conf)
    tp, conf, pred_cls = tp[order], conf[order], pred_cls[order]
    classes, nt = np.unique(target_cls, return_counts=True)
    nc = classes.shape[0]
    px = np.linspace(0, 1, 1000)
    ap, p, r = np.zeros((nc, tp.shape[1])), np.zeros((nc, 1000)), np.zeros((nc, 1000))
    for ci, c in enumerate(classes):
        sel = pred_cls == c
        n_l, n_p = nt[ci], sel.sum()
        if n_p == 0 or n_l == 0:
            continue
        fpc, tpc = (1 - tp[sel]).cumsum(0), tp[sel].cumsum(0)
        recall = tpc / (n_l + eps)
        r[ci] = np.interp(-px, -conf[sel], recall[:, 0], left=0)       # negated: the abscissa must increase
        precision = tpc / (tpc + fpc)
        p[ci] = np.interp(-px, -conf[sel], precision[:, 0], left=1)
        for j in range(tp.shape[1]):
            ap[ci, j] = compute_ap(recall[:, j], precision[:, j])[0]
    f1 = 2 * p * r / (p + r + eps)
    best = _box_filter(f1.mean(0), 0.1).argmax()
    p, r, f1 = p[:, best], r[:, best], f1[:, best]
    tpn = (r * nt).round()
    fpn = (tpn / (p + eps) - tpn).round()
    return tpn, fpn, p, r, f1, ap, classes.astype(int)


class Metric:
    """metrics.py:507-633."""

    def __init__(self):
        self.p, self.r, self.f1, self.all_ap, self.ap_class_index, self.nc = [], [], [], [], [], 0

    ap50 = property(lambda s: s.all_ap[:, 0] if len(s.all_ap) else [])
    ap = property(lambda s: s.all_ap.mean(1) if len(s.all_ap) else [])
    mp = property(lambda s: s.p.mean() if len(s.p) else 0.0)
    mr = property(lambda s: s.r.mean() if len(s.r) else 0.0)
    map50 = property(lambda s: s.all_ap[:, 0].mean() if len(s.all_ap) else 0.0)
    map75 = property(lambda s: s.all_ap[:, 5].mean() if len(s.all_ap) else 0.0)
    map = property(lambda s: s.all_ap.mean() if len(s.all_ap) else 0.0)

    def mean_results(self):
        return [self.mp, self.mr, self.map50, self.map]

    def class_result(self, i):
        return self.p[i], self.r[i], self.ap50[i], self.ap[i]

    @property
    def maps(self):
        maps = np.zeros(self.nc) + self.map
        for i, c in enumerate(self.ap_class_index):
            maps[c] = self.ap[i]
        return maps

    def fitness(self):
        return (np.array(self.mean_results()) * [0.0, 0.0, 0.1, 0.9]).sum()

    def update(self, results):
        self.p, self.r, self.f1, self.all_ap, self.ap_class_index = results


class DetMetrics:
    """metrics.py:635-717 (no plots)."""

    def __init__(self, save_dir=None, plot=False, on_plot=None, names=()):
        self.save_dir, self.plot, self.on_plot, self.names = save_dir, False, on_plot, names
        self.box = Metric()
        self.speed = {"preprocess": 0.0, "inference": 0.0, "loss": 0.0, "postprocess": 0.0}

    def process(self, tp, conf, pred_cls, target_cls):
        self.box.nc = len(self.names)
        self.box.update(ap_per_class(tp, conf, pred_cls, target_cls)[2:])

    keys = property(lambda s: ["metrics/precision(B)", "metrics/recall(B)", "metrics/mAP50(B)", "metrics/mAP50-95(B)"])
    maps = property(lambda s: s.box.maps)
    fitness = property(lambda s: s.box.fitness())
    ap_class_index = property(lambda s: s.box.ap_class_index)

    def mean_results(self):
        return self.box.mean_results()

    def class_result(self, i):
        return self.box.class_result(i)

    @property
    def results_dict(self):
        return dict(zip(self.keys + ["fitness"], self.mean_results() + [self.fitness]))


def _box_iou(a, b, eps=1e-7):
    """box_iou (metrics.py:52-72): (N, 4) x (M, 4) xyxy -> (N, M)."""
    (a1, a2), (b1, b2) = a.unsqueeze(1).chunk(2, 2), b.unsqueeze(0).chunk(2, 2)
    inter = (torch.min(a2, b2) - torch.max(a1, b1)).clamp_(0).prod(2)
    return inter / ((a2 - a1).prod(2) + (b2 - b1).prod(2) - inter + eps)


class ConfusionMatrix:
    """metrics.py:177-290, detection task (rows = predicted class, columns = true class, index nc = background)."""

    def __init__(self, nc, conf=0.25, iou_thres=0.45, task="detect"):
        if task != "detect":
            raise NotImplementedError("ConfusionMatrix: the detection task is the one on this path")
        self.task, self.nc, self.conf, self.iou_thres = task, nc, conf, iou_thres
        self.matrix = np.zeros((nc + 1, nc + 1))

    def process_batch(self, detections, labels):
        if detections is None:
            for gc in labels.int():
                self.matrix[self.nc, gc] += 1
            return
        detections = detections[detections[:, 4] > self.conf]
        gt = labels[:, 0].int().cpu().numpy()
        dc = detections[:, 5].int().cpu().numpy()
        iou = _box_iou(labels[:, 1:].float(), detections[:, :4].float())
        li, di = torch.where(iou > self.iou_thres)
        if li.shape[0]:
            m = torch.cat((torch.stack((li, di), 1).float(), iou[li, di][:, None]), 1).cpu().numpy()
            if li.shape[0] > 1:          # best IoU first, then one match per detection and per label
                m = m[m[:, 2].argsort()[::-1]]
                m = m[np.unique(m[:, 1], return_index=True)[1]]
                m = m[m[:, 2].argsort()[::-1]]
                m = m[np.unique(m[:, 0], return_index=True)[1]]
        else:
            m = np.zeros((0, 3))
        any_match = m.shape[0] > 0
        m0, m1, _ = m.transpose().astype(int)
        for i, gc in enumerate(gt):
            j = m0 == i
            if any_match and j.sum() == 1:
                self.matrix[dc[m1[j]], gc] += 1
            else:
                self.matrix[self.nc, gc] += 1
        if any_match:
            for i, c in enumerate(dc):
                if not (m1 == i).any():
                    self.matrix[c, self.nc] += 1

    def tp_fp(self):
        tp = self.matrix.diagonal()
        fp = self.matrix.sum(1) - tp
        return tp[:-1], fp[:-1]


def get_stats(stats, nc, names=None):
    """val.py:124-131: list of per-image (correct, conf, pred_cls, target_cls) tuples (as `update_metrics` builds them)
    -> (results_dict, DetMetrics, targets per class)."""
    metrics = DetMetrics(names=names if names is not None else {i: str(i) for i in range(nc)})
    cols = [torch.cat(x, 0).cpu().numpy() for x in zip(*stats)] if len(stats) else []
    if len(cols) and cols[0].any():
        metrics.process(*cols)
    nt = np.bincount(cols[-1].astype(int), minlength=nc) if len(cols) else np.zeros(nc, dtype=int)
    return metrics.results_dict, metrics, nt
