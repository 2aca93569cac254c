"""Common machinery of the drop-in modules.

Every class keeps the reference's name, constructor signature, attribute names and state_dict
keys; its parameters live in ordinary torch containers (nn.Conv2d,
nn.BatchNorm2d, ...) that are never *called* -- `forward` packs them once per (dtype, device,
parameter version) into the kernels' layouts and dispatches through the C ABI.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


def _version(t) -> int:
    """In-place update counter of a tensor; inference tensors (parameters created / cast under torch.inference_mode,
    as the reference's predictor and validator do, yolo/engine/model.py:222) do not track one and cannot be updated
    in place either, so their storage pointer identifies them."""
    return 0 if t.is_inference() else t._version


class KernelModule(nn.Module):
    """nn.Module whose forward runs sm_100a kernels (eval mode, CUDA tensors only).  In train mode the call is routed
    to the differentiable torch-operator forward of train_forward.py (batch-statistics BatchNorm, autograd backward):
    the training step's B200-native parts are the criterion and the optimizer side (train.py, csrc/train.cu)."""

    def __call__(self, *args, **kwargs):
        if self.training:
            from .. import train_forward
            return train_forward.run(self, *args, **kwargs)
        return super().__call__(*args, **kwargs)

    def _packed(self, name, dtype, device, tensors, builder):
        """Weights of this module in the kernels' layouts, built once per (name, dtype, device) and rebuilt when a
        source parameter is replaced or updated in place (data_ptr / _version).  Packs for different dtypes / devices
        coexist (an fp32 validation pass does not evict the bf16 pack a captured CUDA graph points at); a replaced
        pack is parked in `_pk_old` until `release_stale_packs()` so that a graph captured over it never reads
        freed memory (Engine additionally pins what it captured and refuses to replay over changed parameters)."""
        cache = self.__dict__.setdefault("_pk", {})
        key = (name, dtype, str(device))
        ver = tuple((t.data_ptr(), _version(t)) for t in tensors)
        hit = cache.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        with torch.no_grad():
            val = builder()
        if hit is not None:
            self.__dict__.setdefault("_pk_old", []).append(hit[1])
        cache[key] = (ver, val)
        return val

    def release_stale_packs(self):
        """Drop packs that were superseded by a weight update (call when no captured graph uses them any more)."""
        self.__dict__.pop("_pk_old", None)

    def _check_mode(self, x):
        t = x[0] if isinstance(x, (list, tuple)) else x
        ops.require_cuda(t, f"{type(self).__name__} input")
        if self.training:   # unreachable through __call__; a direct .forward() in train mode
            raise RuntimeError(f"{type(self).__name__}.forward is the inference kernel path; call the module (train mode "
                               "dispatches to train_forward.py) or switch to .eval()")

    def __getstate__(self):  # packed caches hold device pointers; never pickle them
        d = dict(self.__dict__)
        d.pop("_pk", None)
        d.pop("_pk_old", None)
        return d


def act_name(m) -> str | None:
    if m is None or isinstance(m, nn.Identity):
        return None
    if isinstance(m, nn.SiLU):
        return "silu"
    if isinstance(m, nn.ReLU):
        return "relu"
    if isinstance(m, nn.Sigmoid):
        return "sigmoid"
    if isinstance(m, nn.GELU):
        return "gelu"
    raise NotImplementedError(f"activation {type(m).__name__} has no fused epilogue in mgdt_b200")


def ohwi(w: torch.Tensor, dtype, device, stride: int = 1) -> "ops.PackedConv":
    """(Cout, Cin, kh, kw) -> kernel weight pack: contiguous OHWI (Cout, kh, kw, Cin) in the compute dtype
    for the CUDA-core path plus, for bf16 shapes the tcgen05 path takes, its K-major shared-memory image."""
    t32 = w.detach().to(device=device, dtype=torch.float32).permute(0, 2, 3, 1).contiguous()
    return ops.PackedConv(t32.to(dtype), stride, w32=t32)


def f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()
