"""Train-mode forward of the drop-in modules (row f3, SURVEY.md §7.7 first step).

In eval mode every module dispatches to the sm_100a kernels (BatchNorm folded, activations NHWC, nothing
differentiable).  A training step needs batch-statistics BatchNorm and a backward pass, so in train mode
`KernelModule.__call__` routes here: the same parameters (they live in ordinary nn.Conv2d / nn.BatchNorm2d /
nn.GroupNorm / nn.Linear containers with the reference's state_dict keys) are applied with differentiable torch
operators on the GPU (cuDNN / cuBLAS / torchvision.ops.deform_conv2d), and autograd provides the backward.  The
B200-native parts of the training step are the criterion (csrc/train.cu: assigner + losses + their gradient), the
EMA / clip / SGD updates over one flat bucket and the single flat gradient all-reduce (train.py).

Each function states the reference forward it follows (file:line under the reference root).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = ("run",)


def _conv(m, x):                      # nn/modules/conv.py:36-42
    y = m.conv(x)
    if hasattr(m, "bn"):
        y = m.bn(y)
    return m.act(y)


def _bottleneck(m, x):                # nn/modules/block.py:524-526
    y = m.cv2(m.cv1(x))
    return x + y if m.add else y


def _c2f(m, x):                       # nn/modules/block.py:199-203
    parts = list(m.cv1(x).chunk(2, 1))
    for b in m.m:
        parts.append(b(parts[-1]))
    return m.cv2(torch.cat(parts, 1))


def _spr(m, x):                       # nn/modules/spr_module.py:20-31
    n = x.shape[0]
    s = torch.cat((F.adaptive_avg_pool2d(x, 1).reshape(n, -1, 1, 1), F.adaptive_avg_pool2d(x, 2).reshape(n, -1, 1, 1)), 1)
    return torch.sigmoid(m.fc2(F.relu(m.fc1(s))))


def _mspa_c2f(m, x):                  # nn/modules/block.py:243-287
    g = m.nums
    chunks = x.chunk(g, 1)
    outs, sp = [], None
    for i in range(g):
        sp = chunks[i] if i == 0 else sp + chunks[i]
        if i != g - 1:
            sp = m.convs[i](sp)
            outs.append(sp)
        else:
            for b in m.bottleneck:
                sp = b(sp)
                outs.append(sp)
    feat = m.convs[g - 1](torch.cat(outs, 1))
    n, _, h, w = feat.shape
    gates = torch.cat([m.attention(t) for t in feat.chunk(g, 1)], 1).view(n, g, m.outwidth, 1, 1)
    weighted = feat.view(n, g, m.outwidth, h, w) * torch.softmax(gates, 1)
    return weighted.reshape(n, g * m.outwidth, h, w)


def _sppf(m, x):                      # nn/modules/block.py:147-153
    x = m.cv1(x)
    y1 = m.m(x)
    y2 = m.m(y1)
    return m.cv2(torch.cat((x, y1, y2, m.m(y2)), 1))


def _simfusion4(m, x):                # nn/modules/block.py:294-307
    x_l, x_m, x_s, x_n = x
    size = x_s.shape[2:]
    return torch.cat([F.adaptive_avg_pool2d(x_l, size), F.adaptive_avg_pool2d(x_m, size), x_s,
                      F.interpolate(x_n, size=size, mode="bilinear", align_corners=False)], 1)


def _simfusion3(m, x):                # nn/modules/block.py:318-329
    size = x[1].shape[2:]
    a = m.cv1(F.adaptive_avg_pool2d(x[0], size))
    b = m.cv2(x[1])
    c = m.cv3(F.interpolate(x[2], size=size, mode="bilinear", align_corners=False))
    return m.cv_fuse(torch.cat((a, b, c), 1))


def _convnext(m, x):                  # nn/modules/convnextv2.py:33-45, utils.py:161-182
    t = m.dwconv(x).permute(0, 2, 3, 1)
    t = F.layer_norm(t, m.norm.normalized_shape, m.norm.weight, m.norm.bias, m.norm.eps)
    t = m.act(m.pwconv1(t))
    gx = torch.norm(t, p=2, dim=(1, 2), keepdim=True)
    t = m.grn.gamma * (t * (gx / (gx.mean(dim=-1, keepdim=True) + 1e-6))) + m.grn.beta + t
    return x + m.pwconv2(t).permute(0, 3, 1, 2)


def _ifm(m, x):                       # nn/modules/block.py:340-342
    return m.conv(x)


def _inject(m, x):                    # nn/modules/block.py:368-399
    x_l, x_g = x
    size = x_l.shape[2:]
    info = x_g.split(m.global_inp, dim=1)[m.flag]
    local = m.local_embedding(x_l)
    gact, gfeat = m.global_act(info), m.global_embedding(info)
    if size[0] < x_g.shape[2]:
        gate, gfeat = F.adaptive_avg_pool2d(gact, size), F.adaptive_avg_pool2d(gfeat, size)
    else:
        gate = F.interpolate(F.relu6(gact + 3) / 6, size=size, mode="bilinear", align_corners=False)
        gfeat = F.interpolate(gfeat, size=size, mode="bilinear", align_corners=False)
    return local * gate + gfeat


def _dydcn(m, x, offset, mask):       # nn/modules/block.py:426-432 (mmcv ModulatedDeformConv2d == torchvision deform_conv2d)
    from torchvision.ops import deform_conv2d
    c = m.conv
    y = deform_conv2d(x.contiguous(), offset, c.weight, c.bias, stride=c.stride, padding=c.padding, mask=mask)
    return m.norm(y) if m.with_norm else y


def _conv_gn(m, x):                   # nn/modules/head.py:78-81
    return m.act(m.gn(m.conv(x)))


def _task_decomp(m, feat, avg_feat=None):   # nn/modules/head.py:112-131 (the reduction conv's bias is never applied)
    b, c, h, w = feat.shape
    if avg_feat is None:
        avg_feat = F.adaptive_avg_pool2d(feat, (1, 1))
    att = torch.sigmoid(m.la_conv2(F.relu(m.la_conv1(avg_feat))))
    wdyn = att.reshape(b, 1, m.stacked_convs, 1) * m.reduction_conv.conv.weight.reshape(1, m.feat_channels, m.stacked_convs, m.feat_channels)
    out = torch.bmm(wdyn.reshape(b, m.feat_channels, m.in_channels), feat.reshape(b, m.in_channels, h * w))
    return F.relu(out.reshape(b, m.feat_channels, h, w))


def _detect(m, x):                    # nn/modules/head.py:155-164 (train mode returns the per-level maps)
    for i in range(m.nl):
        x[i] = torch.cat((m.cv2[i](x[i]), m.cv3[i](x[i])), 1)
    return x


def _tood(m, x):                      # nn/modules/head.py:498-533
    for i in range(m.nl):
        stack = [m.share_conv[0](x[i])]
        for layer in m.share_conv[1:]:
            stack.append(layer(stack[-1]))
        feat = torch.cat(stack, 1)
        avg = F.adaptive_avg_pool2d(feat, (1, 1))
        cls_feat, reg_feat = m.cls_decomp(feat, avg), m.reg_decomp(feat, avg)
        om = m.spatial_conv_offset(feat)
        reg_feat = m.DyDCNV2(reg_feat, om[:, :m.offset_dim], om[:, m.offset_dim:].sigmoid())
        prob = m.cls_prob_conv2(F.relu(m.cls_prob_conv1(feat))).sigmoid()
        x[i] = torch.cat((m.cv2(F.relu(reg_feat)), m.cv3(cls_feat * prob)), 1)
    return x


def _concat(m, x):                    # nn/modules/conv.py:294-297
    return torch.cat(list(x), m.d)


def _upsample(m, x):                  # torch.nn.Upsample
    return F.interpolate(x, size=m.size, scale_factor=m.scale_factor, mode=m.mode)


_TABLE = {"Conv": _conv, "Bottleneck": _bottleneck, "C2f": _c2f, "SPRModule": _spr, "MSPA_C2f": _mspa_c2f, "SPPF": _sppf,
          "SimFusion_4in": _simfusion4, "SimFusion_3in": _simfusion3, "ConvNeXtV2_Block": _convnext, "IFM": _ifm,
          "InjectionMultiSum_Auto_pool": _inject, "DyDCNv2": _dydcn, "Conv_GN": _conv_gn, "TaskDecomposition": _task_decomp,
          "Detect": _detect, "TOODHead": _tood, "Concat": _concat, "Upsample": _upsample}


def run(module: nn.Module, *args, **kwargs):
    """Differentiable forward of one drop-in module with the reference's semantics (NCHW tensors)."""
    fn = _TABLE.get(type(module).__name__)
    if fn is None:
        raise NotImplementedError(f"{type(module).__name__}: no train-mode forward")
    extra = {k: v for k, v in kwargs.items() if v is not None}
    if extra and fn not in (_dydcn, _task_decomp):
        raise TypeError(f"{type(module).__name__}: kernel-only arguments {sorted(extra)} in train mode")
    return fn(module, *args, **extra)
