"""Conv / Concat / Upsample -- mirrors of nn/modules/conv.py (reference) on the B200 kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .base import KernelModule, act_name, f32, ohwi

__all__ = ("Conv", "Concat", "Upsample", "autopad")


def autopad(k, p=None, d=1):
    """'same' padding (nn/modules/conv.py:16-22)."""
    if d > 1:
        k = d * (k - 1) + 1 if isinstance(k, int) else [d * (x - 1) + 1 for x in k]
    if p is None:
        p = k // 2 if isinstance(k, int) else [x // 2 for x in k]
    return p


def fold_conv_bn(conv: nn.Conv2d, bn):
    """fuse_conv_and_bn (yolo/utils/torch_utils.py:114-135) in fp32: returns (W', b')."""
    w = conv.weight.detach().float()
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
    if bn is None:
        return w, b
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    return w * scale.view(-1, 1, 1, 1), (b - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()


class Conv(KernelModule):
    """Conv2d(k, s, p='same', bias=False) + BatchNorm2d + activation (nn/modules/conv.py:25-42).
    One fused kernel: BN is folded into the weights at pack time, the activation runs in the
    epilogue.  Extra keyword arguments expose the kernel's fused input/output options to the
    enclosing block (`out` = channel slice of a concat buffer, `residual`, `pre_add`)."""
    default_act = nn.SiLU()

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, d=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, autopad(k, p, d), groups=g, dilation=d, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)  # values initialize_weights sets (torch_utils.py:254-256)
        self.act = self.default_act if act is True else act if isinstance(act, nn.Module) else nn.Identity()

    def _pack(self, dtype, device):
        bn = getattr(self, "bn", None)
        tensors = [self.conv.weight] + ([self.conv.bias] if self.conv.bias is not None else [])
        if bn is not None:
            tensors += [bn.weight, bn.bias, bn.running_mean, bn.running_var]

        def build():
            w, b = fold_conv_bn(self.conv, bn)
            return ohwi(w, dtype, device, self.conv.stride[0]), f32(b, device)

        return self._packed("conv", dtype, device, tensors, build)

    def forward(self, x, out=None, residual=None, pre_add=None, in_scale=None, stat=None):
        self._check_mode(x)
        c = self.conv
        if c.groups != 1 or c.dilation != (1, 1) or c.kernel_size[0] != c.kernel_size[1] or c.stride[0] != c.stride[1]:
            raise NotImplementedError("Conv: only square, dense (groups=1, dilation=1) convolutions are on the hot path")
        x = ops.as_act(x)
        w, b = self._pack(x.dtype, x.device)
        return ops.conv2d(x, w, b, c.kernel_size[0], c.stride[0], c.padding[0], act_name(self.act), out=out,
                          residual=residual, pre_add=pre_add, in_scale=in_scale, stat=stat)

    forward_fuse = forward  # after fuse() the reference swaps forward := forward_fuse (tasks.py:137)

    def can_fuse_preprocess(self):
        c = self.conv
        return (c.kernel_size == (3, 3) and c.stride == (2, 2) and c.padding == (1, 1) and c.groups == 1
                and c.in_channels <= 8)

    def forward_image(self, src, out=None):
        """Layer-0 fast path of the engine: `src` is the raw NCHW uint8 (or float32) batch; preprocessing
        (/255, NCHW->NHWC, bf16) is fused into this conv's loader on the tensor cores."""
        c = self.conv
        bn = getattr(self, "bn", None)
        tensors = [c.weight] + ([c.bias] if c.bias is not None else [])
        if bn is not None:
            tensors += [bn.weight, bn.bias, bn.running_mean, bn.running_var]

        def build():
            w, b = fold_conv_bn(c, bn)                                   # (Cout, Cin, 3, 3)
            k = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)            # k = (dy*3+dx)*Cin + ci
            kp = (k.shape[1] + 15) // 16 * 16
            wp = torch.zeros((w.shape[0], 1, 1, kp), dtype=torch.float32, device=src.device)
            wp[:, 0, 0, :k.shape[1]] = k.to(src.device)
            pk = ops.PackedConv(wp.to(torch.bfloat16), 1, w32=wp)
            if c.in_channels == 3:
                pk.stem_u8 = ops.stem_u8_pack(wp, w.shape[0])            # uint8 sources: warp-MMA stem (csrc/stem_mma.cu)
            return pk, f32(b, src.device)

        pw, b = self._packed("stem", torch.bfloat16, src.device, tensors, build)
        return ops.stem_conv(src, pw, b, c.out_channels, act_name(self.act), out=out)

    def fuse(self):
        """Fold BN into self.conv the way BaseModel.fuse does (nn/tasks.py:133-137)."""
        if hasattr(self, "bn"):
            w, b = fold_conv_bn(self.conv, self.bn)
            c = self.conv
            fused = nn.Conv2d(c.in_channels, c.out_channels, c.kernel_size, c.stride, c.padding, bias=True)
            fused = fused.requires_grad_(False).to(c.weight.device)
            fused.weight.copy_(w)
            fused.bias.copy_(b)
            self.conv = fused
            delattr(self, "bn")
        return self


class Concat(KernelModule):
    """torch.cat(x, dim) (nn/modules/conv.py:287-297); channel concat of NHWC maps is a strided copy
    of each input into its slice."""

    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension

    def forward(self, x):
        self._check_mode(x)
        if self.d != 1:
            raise NotImplementedError("Concat: only channel concatenation (dimension=1) is on the hot path")
        xs = [t if isinstance(t, DeferredResample) else ops.as_act(t) for t in x]
        n, _, h, w = xs[0].shape
        ctot = sum(t.shape[1] for t in xs)
        out = ops.new_act(n, ctot, h, w, xs[0].dtype, xs[0].device)
        c0 = 0
        for t in xs:
            c = t.shape[1]
            if isinstance(t, DeferredResample):
                ops.resample(t.src, h, w, t.mode, out=out[:, c0:c0 + c])
            else:
                ops.resample(t, h, w, ops.RS_COPY, out=out[:, c0:c0 + c])
            c0 += c
        return out


class DeferredResample:
    """An un-materialised nn.Upsample output; Concat resamples straight into its slice."""

    def __init__(self, src, h, w, mode):
        self.src, self.mode = src, mode
        self.shape = (src.shape[0], src.shape[1], h, w)
        self.dtype, self.device, self.is_cuda = src.dtype, src.device, True

    def materialize(self):
        return ops.resample(self.src, self.shape[2], self.shape[3], self.mode)


class Upsample(nn.Upsample):
    """nn.Upsample(None, 2, 'nearest') of the PAN neck (models/v8/yolov8.yaml:31,35)."""

    defer = False  # set by DetectionModel when the only consumer is the next Concat

    def forward(self, x):
        if self.training:
            return nn.Upsample.forward(self, x)
        ops.require_cuda(x, "Upsample input")
        if self.mode not in ("nearest", "bilinear") or self.scale_factor is None:
            raise NotImplementedError("Upsample: nearest/bilinear with scale_factor only")
        if self.mode == "bilinear" and self.align_corners:
            raise NotImplementedError("Upsample: align_corners=True is not on the hot path")
        x = ops.as_act(x)
        sf = self.scale_factor
        sh, sw = (sf, sf) if not isinstance(sf, (tuple, list)) else sf
        h, w = int(x.shape[2] * sh), int(x.shape[3] * sw)
        mode = ops.RS_NEAREST if self.mode == "nearest" else ops.RS_BILINEAR
        if self.defer:
            return DeferredResample(x, h, w, mode)
        return ops.resample(x, h, w, mode)
