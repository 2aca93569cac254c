"""Deterministic synthetic weights and inputs.

There is no network, so every parity test, golden fixture and bench line uses
random-init weights of the real architecture plus synthetic 640x640 inputs.
The recipe is keyed on the *state_dict key name*, not on construction order, so
the live reference model (oracle/make_golden.py, this container only), the
oracle restatement and the B200 modules all receive bit-identical tensors as
long as their state_dict keys and shapes agree -- which is itself part of the
drop-in boundary.

Normalisation state is randomised on purpose: the reference defaults make
BatchNorm ~identity and GRN a no-op (gamma = beta = 0, nn/modules/utils.py:176-177),
which would hide bugs.
"""
from __future__ import annotations

import math
import zlib

import torch

__all__ = ["synth_state_dict", "synth_images", "synth_predictions", "raise_cls_bias"]


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _normal(shape, std, g):
    return torch.randn(shape, generator=g, dtype=torch.float32) * std


def _uniform(shape, lo, hi, g):
    return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo


def synth_state_dict(template: dict, seed: int = 1) -> dict:
    """Return a new state_dict with the same keys/shapes/dtypes as `template`.

    Rules (by key suffix):
      *.num_batches_tracked, *.dfl.conv.weight, *.scale.*.scale  -> kept as is
      norm scale  (bn/gn/norm .weight, 1-D)                      -> U(0.5, 1.5)
      norm shift / any bias                                      -> N(0, 0.1)
      running_mean                                               -> N(0, 0.1)
      running_var                                                -> U(0.5, 1.5)
      grn.gamma / grn.beta                                       -> N(0, 0.1)
      conv / linear weights                                      -> N(0, sqrt(2 / fan_in))
    """
    out = {}
    for key, t in template.items():
        if key.endswith("num_batches_tracked") or key.endswith("dfl.conv.weight") or ".scale." in key:
            out[key] = t.clone()
            continue
        g = _gen(key, seed)
        shape = tuple(t.shape)
        if key.endswith("running_var"):
            v = _uniform(shape, 0.5, 1.5, g)
        elif key.endswith("running_mean"):
            v = _normal(shape, 0.1, g)
        elif key.endswith("grn.gamma") or key.endswith("grn.beta"):
            v = _normal(shape, 0.1, g)
        elif key.endswith(".bias"):
            v = _normal(shape, 0.1, g)
        elif key.endswith(".weight") and t.dim() == 1:
            v = _uniform(shape, 0.5, 1.5, g)
        elif key.endswith(".weight"):
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            v = _normal(shape, math.sqrt(2.0 / max(fan_in, 1)), g)
        else:  # unknown tensor kind: small noise, never silently constant
            v = _normal(shape, 0.1, g)
        out[key] = v.to(t.dtype)
    return out


def raise_cls_bias(sd: dict, value: float = -2.0) -> dict:
    """Random-init class biases never pass `conf` (head.py:186,568), so for the
    end-to-end NMS variant lift the final class-predictor bias."""
    for key in sd:
        if key.endswith("cv3.bias") or (".cv3." in key and key.endswith(".2.bias")):
            sd[key] = torch.full_like(sd[key], value)
    return sd


def synth_images(batch: int, size: int = 640, seed: int = 0, ch: int = 3, h: int | None = None,
                 w: int | None = None) -> torch.Tensor:
    """U[0,1) fp32 (B, ch, H, W), as the reference's predictor feeds /255 floats
    (yolo/engine/predictor.py:127-129)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.rand((batch, ch, h or size, w or size), generator=g, dtype=torch.float32)


def synth_predictions(batch: int, nc: int, anchors: int, seed: int = 2, img: float = 640.0) -> torch.Tensor:
    """Synthetic decode output (B, 4+nc, A) for NMS: centres U(0,img), w/h
    log-uniform 8..256 px, scores heavy near 0 and tie-free."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    cx = torch.rand((batch, 1, anchors), generator=g) * img
    cy = torch.rand((batch, 1, anchors), generator=g) * img
    lw = torch.rand((batch, 2, anchors), generator=g) * (math.log(256.0) - math.log(8.0)) + math.log(8.0)
    wh = lw.exp()
    # Tie-free by construction: a random permutation of n distinct levels pushed through the
    # monotone map t -> t^8 (mass near 0, thin tail to 1, like Beta(0.5, 4)).  The reference's
    # argsort (ops.py:244) is unstable, so golden vectors must not contain equal scores.
    n = batch * nc * anchors
    perm = torch.randperm(n, generator=g).to(torch.float64)
    sc = (((perm + 0.5) / n) ** 8).to(torch.float32).reshape(batch, nc, anchors)
    return torch.cat((cx, cy, wh, sc), 1).contiguous()
