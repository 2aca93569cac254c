"""Pre-processing on the device: LetterBox + BasePredictor.preprocess.

Mirrors `ultralytics.yolo.data.augment.LetterBox` (augment.py:538-593: same constructor, same geometry) and the
list branch of `BasePredictor.preprocess` (predictor.py:121-125: stack, BGR->RGB, BHWC->BCHW).  The images stay
uint8 -- the /255 and the bf16 cast are fused into the stem convolution -- and every byte equals what the
reference's cv2.resize / copyMakeBorder pipeline produces (tests/test_gpu_prepost.py, tests/golden/prepost.npz).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops

__all__ = ("LetterBox", "preprocess_images", "letterbox_geometry")


def letterbox_geometry(shape, new_shape=(640, 640), auto=False, scaleFill=False, scaleup=True, stride=32):
    """(new_unpad (w, h), (top, bottom, left, right), ratio (w, h), (dw, dh)) exactly as LetterBox.__call__ computes
    them (augment.py:554-583)."""
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
    if not scaleup:
        r = min(r, 1.0)
    ratio = r, r
    new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = np.mod(dw, stride), np.mod(dh, stride)
    elif scaleFill:
        dw, dh = 0.0, 0.0
        new_unpad = (new_shape[1], new_shape[0])
        ratio = new_shape[1] / shape[1], new_shape[0] / shape[0]
    dw /= 2
    dh /= 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return new_unpad, (top, bottom, left, right), ratio, (dw, dh)


def _to_device_image(im, device):
    if isinstance(im, np.ndarray):
        im = torch.from_numpy(np.ascontiguousarray(im))
    if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3:
        raise ValueError("expected an (h, w, 3) uint8 image")
    return im.to(device, non_blocking=True).contiguous()


class LetterBox:
    """Resize image and padding for detection -- same constructor and `(labels=None, image=None)` call as the
    reference's LetterBox; `image` is an (h, w, 3) uint8 array / tensor, the result an (H, W, 3) uint8 CUDA tensor.
    The `labels` path (training-time label rescaling) is not on the inference hot path."""

    def __init__(self, new_shape=(640, 640), auto=False, scaleFill=False, scaleup=True, stride=32, device="cuda"):
        self.new_shape = new_shape
        self.auto = auto
        self.scaleFill = scaleFill
        self.scaleup = scaleup
        self.stride = stride
        self.device = torch.device(device)

    def __call__(self, labels=None, image=None):
        if labels:
            raise NotImplementedError("LetterBox: label rescaling belongs to the training data pipeline")
        if image is None:
            raise ValueError("LetterBox: image is required")
        img = _to_device_image(image, self.device)
        new_unpad, (top, bottom, left, right), _, _ = letterbox_geometry(img.shape[:2], self.new_shape, self.auto,
                                                                         self.scaleFill, self.scaleup, self.stride)
        out = torch.empty((new_unpad[1] + top + bottom, new_unpad[0] + left + right, 3), dtype=torch.uint8, device=img.device)
        return ops.letterbox_u8(img, out, (new_unpad[1], new_unpad[0]), (top, left), swap_rb=False, out_hwc=True)


def preprocess_images(ims, new_shape=(640, 640), auto=False, stride=32, device="cuda", out=None):
    """List of (h, w, 3) BGR uint8 images -> (N, 3, H, W) RGB uint8 CUDA batch, one launch per image writing straight
    into its NCHW slot (letterbox + channel swap + transpose fused).  All images must letterbox to the same (H, W)
    (always true with auto=False; with auto=True the reference requires equal source shapes, predictor.py:138-139).
    Returns (batch, metas) with metas[i] = (orig_shape, ratio, (dw, dh)) for scale_boxes."""
    device = torch.device(device)
    geo = [letterbox_geometry(im.shape[:2], new_shape, auto, False, True, stride) for im in ims]
    hw = {(g[0][1] + g[1][0] + g[1][1], g[0][0] + g[1][2] + g[1][3]) for g in geo}
    if len(hw) != 1:
        raise ValueError(f"preprocess_images: images letterbox to different shapes {sorted(hw)}")
    H, W = hw.pop()
    if out is None:
        out = torch.empty((len(ims), 3, H, W), dtype=torch.uint8, device=device)
    elif tuple(out.shape) != (len(ims), 3, H, W) or out.dtype != torch.uint8:
        raise ValueError("preprocess_images: `out` has the wrong shape / dtype")
    metas = []
    for i, (im, (new_unpad, (top, bottom, left, right), ratio, dwdh)) in enumerate(zip(ims, geo)):
        img = _to_device_image(im, device)
        ops.letterbox_u8(img, out[i], (new_unpad[1], new_unpad[0]), (top, left), swap_rb=True, out_hwc=False)
        metas.append((tuple(int(v) for v in im.shape[:2]), ratio, dwdh))
    return out, metas
