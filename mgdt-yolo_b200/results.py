"""Results / Boxes -- mirrors of yolo/engine/results.py (reference :20-64 BaseTensor, :66-180 Results, :354-440 Boxes)
for the detection task, on device tensors.

The containers keep the reference's attribute and property names.  Coordinate conversions of CUDA data run through
mgdt_box_convert (bit-exact with the reference's fp32 arithmetic); data moved to the host with .cpu() / .numpy() is a
plain copy and its views use the reference's own expressions (container behaviour, not the hot path).
"""
from __future__ import annotations

from copy import deepcopy

import numpy as np
import torch

from . import ops

__all__ = ("BaseTensor", "Boxes", "Results", "build_results")


class BaseTensor:
    """Base tensor class with device handling (results.py:20-64)."""

    def __init__(self, data, orig_shape) -> None:
        assert isinstance(data, (torch.Tensor, np.ndarray))
        self.data = data
        self.orig_shape = orig_shape

    @property
    def shape(self):
        return self.data.shape

    def cpu(self):
        return self if isinstance(self.data, np.ndarray) else self.__class__(self.data.cpu(), self.orig_shape)

    def numpy(self):
        return self if isinstance(self.data, np.ndarray) else self.__class__(self.data.cpu().numpy(), self.orig_shape)

    def cuda(self):
        return self.__class__(torch.as_tensor(self.data).cuda(), self.orig_shape)

    def to(self, *args, **kwargs):
        return self.__class__(torch.as_tensor(self.data).to(*args, **kwargs), self.orig_shape)

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        return self.__class__(self.data[idx], self.orig_shape)


def _xyxy2xywh_host(x):
    y = x.clone() if isinstance(x, torch.Tensor) else np.copy(x)
    y[..., 0] = (x[..., 0] + x[..., 2]) / 2
    y[..., 1] = (x[..., 1] + x[..., 3]) / 2
    y[..., 2] = x[..., 2] - x[..., 0]
    y[..., 3] = x[..., 3] - x[..., 1]
    return y


class Boxes(BaseTensor):
    """Detection boxes (n, 6) = xyxy, conf, cls, or (n, 7) with a track id before conf (results.py:354-440)."""

    def __init__(self, boxes, orig_shape) -> None:
        if boxes.ndim == 1:
            boxes = boxes[None, :]
        n = boxes.shape[-1]
        assert n in (6, 7), f'expected `n` in [6, 7], but got {n}'  # xyxy, (track_id), conf, cls
        super().__init__(boxes, orig_shape)
        self.is_track = n == 7
        self.orig_shape = orig_shape
        self._views = {}

    @property
    def xyxy(self):
        return self.data[:, :4]

    @property
    def conf(self):
        return self.data[:, -2]

    @property
    def cls(self):
        return self.data[:, -1]

    @property
    def id(self):
        return self.data[:, -3] if self.is_track else None

    def _view(self, mode):
        v = self._views.get(mode)
        if v is None:
            d = self.data
            if isinstance(d, torch.Tensor) and d.is_cuda and d.dtype == torch.float32:
                v = ops.box_convert(d if d.stride(-1) == 1 else d.contiguous(), mode, self.orig_shape[1], self.orig_shape[0])
            else:  # host copy: the reference's expressions (results.py:405-430)
                v = _xyxy2xywh_host(self.xyxy) if mode & 1 else (self.xyxy.clone() if isinstance(d, torch.Tensor) else np.copy(self.xyxy))
                if mode & 2:
                    v[..., [0, 2]] /= self.orig_shape[1]
                    v[..., [1, 3]] /= self.orig_shape[0]
            self._views[mode] = v
        return v

    @property
    def xywh(self):
        return self._view(1)

    @property
    def xyxyn(self):
        return self._view(2)

    @property
    def xywhn(self):
        return self._view(3)

    @property
    def boxes(self):
        """The raw tensor (deprecated in the reference, results.py:432-435)."""
        return self.data


class Results:
    """Inference results of one image (results.py:66-180), detection fields only: orig_img, orig_shape, boxes, names,
    path, speed.  Masks / keypoints / probs belong to other tasks and stay None."""

    def __init__(self, orig_img, path, names, boxes=None, masks=None, probs=None, keypoints=None) -> None:
        if masks is not None or probs is not None or keypoints is not None:
            raise NotImplementedError("Results: masks / probs / keypoints belong to the segmentation, classification and pose tasks")
        self.orig_img = orig_img
        self.orig_shape = orig_img.shape[:2]
        self.boxes = Boxes(boxes, self.orig_shape) if boxes is not None else None
        self.masks = None
        self.probs = None
        self.keypoints = None
        self.speed = {'preprocess': None, 'inference': None, 'postprocess': None}
        self.names = names
        self.path = path
        self._keys = ('boxes', 'masks', 'probs', 'keypoints')

    def keys(self):
        return [k for k in self._keys if getattr(self, k) is not None]

    def new(self):
        return Results(orig_img=self.orig_img, path=self.path, names=self.names)

    def _apply(self, fn):
        r = self.new()
        for k in self.keys():
            setattr(r, k, fn(getattr(self, k)))
        return r

    def __getitem__(self, idx):
        return self._apply(lambda v: v[idx])

    def update(self, boxes=None, masks=None, probs=None):
        if boxes is not None:
            self.boxes = Boxes(boxes, self.orig_shape)

    def cpu(self):
        return self._apply(lambda v: v.cpu())

    def numpy(self):
        return self._apply(lambda v: v.numpy())

    def cuda(self):
        return self._apply(lambda v: v.cuda())

    def to(self, *args, **kwargs):
        return self._apply(lambda v: v.to(*args, **kwargs))

    def __len__(self):
        for k in self.keys():
            return len(getattr(self, k))
        return 0

    def verbose(self):
        """Per-class detection counts as the reference logs them (results.py:206-224)."""
        boxes = self.boxes
        if boxes is None or len(boxes) == 0:
            return '(no detections), '
        s = ''
        cls = boxes.cls if not isinstance(boxes.cls, torch.Tensor) else boxes.cls.cpu()
        for c in np.unique(np.asarray(cls)):
            n = int((np.asarray(cls) == c).sum())
            s += f"{n} {self.names[int(c)]}{'s' * (n > 1)}, "
        return s

    def __deepcopy__(self, memo):
        r = self.new()
        r.boxes = None if self.boxes is None else Boxes(deepcopy(self.boxes.data, memo), self.orig_shape)
        return r


def build_results(dets, orig_imgs, paths, names):
    """DetectionPredictor.postprocess's tail (yolo/v8/detect/predict.py:17-30): one Results per image from the
    (already scale_boxes'ed) detections."""
    out = []
    for i, d in enumerate(dets):
        img = orig_imgs[i] if isinstance(orig_imgs, (list, tuple)) else orig_imgs
        out.append(Results(orig_img=img, path=paths[i] if paths else None, names=names, boxes=d))
    return out
