"""Detection heads -- mirrors of nn/modules/head.py (reference): Detect, TOODHead and their parts."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import ops
from .base import KernelModule, act_name, f32, ohwi
from .block import DFL, DyDCNv2
from .conv import Conv, autopad

__all__ = ("Conv_GN", "TaskDecomposition", "Detect", "TOODHead", "Scale")


class Scale(nn.Module):
    """Learnable scalar (mmcv.cnn.Scale); constructed by TOODHead (head.py:495) but never applied (:527)."""

    def __init__(self, scale=1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(scale, dtype=torch.float))


class _ConvHolder(nn.Module):
    """Parameter container standing where mmcv.cnn.ConvModule stands (head.py:96-104): `.conv` with a bias."""

    def __init__(self, cin, cout, k, bias=True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, bias=bias)
        self.activate = nn.ReLU(inplace=True)


def _pack_plain(conv: nn.Conv2d, dtype, device):
    return ohwi(conv.weight, dtype, device), (f32(conv.bias, device) if conv.bias is not None else None)


class Conv_GN(KernelModule):
    """Conv2d + GroupNorm(16) + SiLU (nn/modules/head.py:67-81): conv kernel, per-(n,c) statistics,
    then one normalise+activate pass."""
    default_act = nn.SiLU()

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, d=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, autopad(k, p, d), groups=g, dilation=d, bias=False)
        self.gn = nn.GroupNorm(16, c2)
        self.act = self.default_act if act is True else act if isinstance(act, nn.Module) else nn.Identity()

    def _pack(self, dtype, device):
        t = [self.conv.weight, self.gn.weight, self.gn.bias]
        return self._packed("cgn", dtype, device, t, lambda: (ohwi(self.conv.weight, dtype, device, self.conv.stride[0]),
                                                              f32(self.gn.weight, device), f32(self.gn.bias, device)))

    def forward(self, x, out=None):
        self._check_mode(x)
        x = ops.as_act(x)
        w, gw, gb = self._pack(x.dtype, x.device)
        c = self.conv
        req = ops.StatReq(1, True)
        raw = ops.conv2d(x, w, None, c.kernel_size[0], c.stride[0], c.padding[0], stat=req)
        a, b = ops.finish_gn(req, raw, self.gn.num_groups, self.gn.eps, gw, gb)   # GroupNorm statistics from the conv's epilogue
        return ops.affine_act(raw, a, b, act=act_name(self.act), out=out)


class TaskDecomposition(KernelModule):
    """Layer-attention task decomposition (nn/modules/head.py:83-131).  The dynamic 1x1 is
    W[o, s*fc + i] * att[b, s]; applying att to the INPUT channels gives the same product, so it is
    the conv kernel's per-(n,c) input scale.  The reduction conv's bias is never applied
    (head.py:122-126)."""

    def __init__(self, feat_channels, stacked_convs, la_down_rate=8, conv_cfg=None, norm_cfg=None):
        super().__init__()
        if norm_cfg is not None:
            raise NotImplementedError("TaskDecomposition: norm_cfg is None everywhere in the reference (head.py:488-489)")
        self.feat_channels = feat_channels
        self.stacked_convs = stacked_convs
        self.in_channels = self.feat_channels * self.stacked_convs
        self.conv_cfg = conv_cfg
        self.norm_cfg = norm_cfg
        self.la_conv1 = nn.Conv2d(self.in_channels, self.in_channels // la_down_rate, 1)
        self.relu = nn.ReLU(inplace=True)
        self.la_conv2 = nn.Conv2d(self.in_channels // la_down_rate, self.stacked_convs, 1, padding=0)
        self.sigmoid = nn.Sigmoid()
        self.reduction_conv = _ConvHolder(self.in_channels, self.feat_channels, 1, bias=True)

    def _tensors(self):
        return [self.la_conv1.weight, self.la_conv1.bias, self.la_conv2.weight, self.la_conv2.bias,
                self.reduction_conv.conv.weight]

    def _pack(self, dtype, device):
        def build():
            return dict(w1=f32(self.la_conv1.weight.flatten(1), device).unsqueeze(0).contiguous(),
                        b1=f32(self.la_conv1.bias, device).unsqueeze(0).contiguous(),
                        w2=f32(self.la_conv2.weight.flatten(1), device).unsqueeze(0).contiguous(),
                        b2=f32(self.la_conv2.bias, device).unsqueeze(0).contiguous(),
                        wr=ohwi(self.reduction_conv.conv.weight, dtype, device))

        return self._packed("td", dtype, device, self._tensors(), build)

    def forward(self, feat, avg_feat=None):
        self._check_mode(feat)
        feat = ops.as_act(feat)
        n, c, h, w = feat.shape
        p = self._pack(feat.dtype, feat.device)
        if avg_feat is None:
            s, _ = ops.chan_stats(feat)
            hw = h * w
        else:
            s, hw = avg_feat.detach().float().reshape(n, 1, c).contiguous(), 1
        att = ops.td_attn(s, hw, p["w1"], p["b1"], p["w2"], p["b2"], self.stacked_convs)
        return ops.conv2d(feat, p["wr"], None, 1, act="relu", in_scale=att[0])


class _HeadBase(KernelModule):
    dynamic = False
    export = False
    shape = None
    anchors = torch.empty(0)
    strides = torch.empty(0)

    def _strides(self):
        """self.stride as python floats, read back once per tensor (a .tolist() on a CUDA tensor is a
        sync and would invalidate a CUDA-graph capture)."""
        st = self.stride
        key = (st.data_ptr(), 0 if st.is_inference() else st._version, str(st.device))
        c = self.__dict__.get("_stride_cache")
        if c is None or c[0] != key:
            c = (key, [float(v) for v in st.tolist()])
            self.__dict__["_stride_cache"] = c
        return c[1]

    def _finish(self, x):
        """Inference tail shared by both heads (head.py:165-177,536-559): DFL + dist2bbox + sigmoid."""
        if self.training:
            return x
        self.shape = x[0].shape
        y = ops.decode(x, self._strides(), self.reg_max, self.nc)
        return y if self.export else (y, x)


class Detect(_HeadBase):
    """YOLOv8 Detect head (nn/modules/head.py:133-186); reg_max = 4 in this fork (head.py:145)."""

    def __init__(self, nc=80, ch=()):
        super().__init__()
        self.nc = nc
        self.nl = len(ch)
        self.reg_max = 4
        self.no = nc + self.reg_max * 4
        self.stride = torch.zeros(self.nl)
        c2, c3 = max((16, ch[0] // 4, self.reg_max * 4)), max(ch[0], self.nc)
        self.cv2 = nn.ModuleList(
            nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 4 * self.reg_max, 1)) for x in ch)
        self.cv3 = nn.ModuleList(nn.Sequential(Conv(x, c3, 3), Conv(c3, c3, 3), nn.Conv2d(c3, self.nc, 1)) for x in ch)
        self.dfl = DFL(self.reg_max) if self.reg_max > 1 else nn.Identity()

    def _pack(self, dtype, device):
        convs = [m[2] for m in self.cv2] + [m[2] for m in self.cv3]
        t = [c.weight for c in convs] + [c.bias for c in convs]
        return self._packed("tails", dtype, device, t, lambda: [_pack_plain(c, dtype, device) for c in convs])

    def forward(self, x):
        self._check_mode(x)
        tails = self._pack(ops.as_act(x[0]).dtype, x[0].device)
        for i in range(self.nl):
            xi = ops.as_act(x[i])
            n, _, h, w = xi.shape
            raw = ops.new_act(n, self.no, h, w, xi.dtype, xi.device)
            wb, bb = tails[i]
            wc, bc = tails[self.nl + i]
            ops.conv2d(self.cv2[i][1](self.cv2[i][0](xi)), wb, bb, 1, out=raw[:, :4 * self.reg_max])
            ops.conv2d(self.cv3[i][1](self.cv3[i][0](xi)), wc, bc, 1, out=raw[:, 4 * self.reg_max:])
            x[i] = raw  # the reference overwrites the caller's list in place (head.py:160)
        return self._finish(x)

    def bias_init(self):
        """head.py:179-186."""
        for a, b, s in zip(self.cv2, self.cv3, self.stride):
            a[-1].bias.data[:] = 1.0
            b[-1].bias.data[:self.nc] = math.log(5 / self.nc / (640 / s) ** 2)


class TOODHead(_HeadBase):
    """Task-aligned dynamic head (nn/modules/head.py:466-572); reg_max = 16 (head.py:481)."""

    def __init__(self, nc, hidc, ch=()):
        super().__init__()
        self.nc = nc
        self.nl = len(ch)
        self.reg_max = 16
        self.no = nc + self.reg_max * 4
        self.stride = torch.zeros(self.nl)
        self.share_conv = nn.Sequential(Conv_GN(hidc, hidc // 2, 3), Conv_GN(hidc // 2, hidc // 2, 3))
        self.cls_decomp = TaskDecomposition(hidc // 2, 2, 16)
        self.reg_decomp = TaskDecomposition(hidc // 2, 2, 16)
        self.DyDCNV2 = DyDCNv2(hidc // 2, hidc // 2)
        self.spatial_conv_offset = nn.Conv2d(hidc, 3 * 3 * 3, 3, padding=1)
        self.offset_dim = 2 * 3 * 3
        self.cls_prob_conv1 = nn.Conv2d(hidc, hidc // 4, 1)
        self.cls_prob_conv2 = nn.Conv2d(hidc // 4, 1, 3, padding=1)
        self.cv2 = nn.Conv2d(hidc // 2, 4 * self.reg_max, 1)
        self.cv3 = nn.Conv2d(hidc // 2, self.nc, 1)
        self.scale = nn.ModuleList(Scale(1.0) for x in ch)
        self.dfl = DFL(self.reg_max) if self.reg_max > 1 else nn.Identity()

    def _pack(self, dtype, device):
        plain = [self.spatial_conv_offset, self.cls_prob_conv1, self.cls_prob_conv2, self.cv2, self.cv3]
        t = [c.weight for c in plain] + [c.bias for c in plain] + self.cls_decomp._tensors() + self.reg_decomp._tensors()  # noqa: E501

        def build():
            d = dict(zip(("off", "p1", "p2", "cv2", "cv3"), (_pack_plain(c, dtype, device) for c in plain)))
            cd, rd = self.cls_decomp, self.reg_decomp
            stack = lambda a, b: torch.stack((f32(a, device), f32(b, device))).contiguous()  # noqa: E731
            d["w1"] = stack(cd.la_conv1.weight.flatten(1), rd.la_conv1.weight.flatten(1))
            d["b1"] = stack(cd.la_conv1.bias, rd.la_conv1.bias)
            d["w2"] = stack(cd.la_conv2.weight.flatten(1), rd.la_conv2.weight.flatten(1))
            d["b2"] = stack(cd.la_conv2.bias, rd.la_conv2.bias)
            d["wcls"] = ohwi(cd.reduction_conv.conv.weight, dtype, device)
            d["wreg"] = ohwi(rd.reduction_conv.conv.weight, dtype, device)
            # the three 1x1 convs that read `feat` (both reduction convs, applied without their bias, head.py:122-126,
            # and cls_prob_conv1; all followed by ReLU) stacked along the output channels: ONE GEMM with per-image weights
            wc, wr, wp = cd.reduction_conv.conv.weight, rd.reduction_conv.conv.weight, self.cls_prob_conv1.weight
            if wc.shape == wr.shape and wc.shape[1] == wp.shape[1] and wc.shape[2:] == (1, 1) and wp.shape[2:] == (1, 1):
                d["w3"] = ohwi(torch.cat([wc, wr, wp], 0), dtype, device)
                d["b3"] = f32(torch.cat([torch.zeros(wc.shape[0] + wr.shape[0], device=wp.device), self.cls_prob_conv1.bias.detach().float()]), device)
            return d

        return self._packed("tood", dtype, device, t, build)

    def forward(self, x):
        self._check_mode(x)
        rm4 = 4 * self.reg_max
        for i in range(self.nl):
            xi = ops.as_act(x[i])
            p = self._pack(xi.dtype, xi.device)
            n, hidc, h, w = xi.shape
            h2 = hidc // 2
            feat = ops.new_act(n, hidc, h, w, xi.dtype, xi.device)      # cat(stack_res_list) (head.py:504-506)
            self.share_conv[0](xi, out=feat[:, :h2])
            self.share_conv[1](feat[:, :h2], out=feat[:, h2:])
            s, _ = ops.chan_stats(feat)                                  # adaptive_avg_pool2d(feat, 1) (:509)
            att = ops.td_attn(s, h * w, p["w1"], p["b1"], p["w2"], p["b2"], 2)
            fused = None
            if "w3" in p and ops.FUSE_TOOD_SIBLINGS:
                fused = ops.conv2d(feat, p["w3"], p["b3"], 1, act="relu", in_scale=att, scale_group_cols=h2)
            if fused is not None:
                cls_feat, reg_feat, prob = fused[:, :h2], fused[:, h2:2 * h2], fused[:, 2 * h2:]
            else:
                cls_feat = ops.conv2d(feat, p["wcls"], None, 1, act="relu", in_scale=att[0])
                reg_feat = ops.conv2d(feat, p["wreg"], None, 1, act="relu", in_scale=att[1])
                prob = ops.conv2d(feat, p["p1"][0], p["p1"][1], 1, act="relu")
            om = ops.conv2d(feat, p["off"][0], p["off"][1], 3)           # offsets 0..17, mask logits 18..26 (:514-517)
            reg = self.DyDCNV2(reg_feat, om[:, :self.offset_dim], om[:, self.offset_dim:], mask_is_logit=True,
                               act="relu")                               # GN, then the F.relu of :528
            raw = ops.new_act(n, self.no, h, w, xi.dtype, xi.device)
            ops.conv2d(reg, p["cv2"][0], p["cv2"][1], 1, out=raw[:, :rm4])
            # cv3(cls_feat * sigmoid(cls_prob_conv2(prob))): one launch, or the two convolutions
            if ops.tood_cls(prob, p["p2"][0].ohwi, p["p2"][1], cls_feat, p["cv3"][0].ohwi, p["cv3"][1], raw[:, rm4:]) is None:
                prob = ops.conv2d(prob, p["p2"][0], p["p2"][1], 3, act="sigmoid")
                ops.conv2d(cls_feat, p["cv3"][0], p["cv3"][1], 1, pix_scale=prob, out=raw[:, rm4:])
            x[i] = raw
        return self._finish(x)

    def bias_init(self):
        """head.py:561-568 (fixed stride 16 in the formula although the level is stride 8)."""
        self.cv2.bias.data[:] = 1.0
        self.cv3.bias.data[:self.nc] = math.log(5 / self.nc / (640 / 16) ** 2)

    def decode_bboxes(self, bboxes):
        raise NotImplementedError("decode is fused into TOODHead.forward (mgdt_decode)")
