"""Block modules -- mirrors of nn/modules/{block,spr_module,convnextv2,utils}.py (reference).

Concatenations never materialise: producers write into channel slices of one NHWC buffer.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .base import KernelModule, act_name, f32, ohwi
from .conv import Conv, fold_conv_bn

__all__ = ("DFL", "SPPF", "Bottleneck", "C2f", "MSPA_C2f", "SPRModule", "SimFusion_4in", "SimFusion_3in", "IFM",
           "ConvNeXtV2_Block", "LayerNorm", "GRN", "h_sigmoid", "InjectionMultiSum_Auto_pool", "DyDCNv2")


def _pk1x1(w32_ohwi, dtype, device):
    """PackedConv of a (Cout,1,1,K) fp32 weight matrix."""
    w32 = w32_ohwi.to(device=device, dtype=torch.float32).contiguous()
    return ops.PackedConv(w32.to(dtype), 1, w32=w32)


class DFL(KernelModule):
    """Distribution-focal-loss integral (nn/modules/block.py:36-54): softmax over c1 bins and
    expectation with the fixed weights 0..c1-1.  Inside the heads it is fused into mgdt_decode."""

    def __init__(self, c1=16):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(c1, dtype=torch.float).view(1, c1, 1, 1)
        self.c1 = c1

    def forward(self, x):
        self._check_mode(x)
        b, c, a = x.shape  # (batch, 4*c1, anchors)
        raw = ops.as_act(x.reshape(b, c, 1, a))
        return ops.decode([raw], [1.0], self.c1, 0, dist_only=True)


class Bottleneck(KernelModule):
    """x + cv2(cv1(x)) (nn/modules/block.py:514-526); the residual add is cv2's epilogue."""

    def __init__(self, c1, c2, shortcut=True, g=1, k=(3, 3), e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, k[0], 1)
        self.cv2 = Conv(c_, c2, k[1], 1, g=g)
        self.add = shortcut and c1 == c2

    def forward(self, x, out=None):
        self._check_mode(x)
        x = ops.as_act(x)
        return self.cv2(self.cv1(x), out=out, residual=x if self.add else None)


class C2f(KernelModule):
    """CSP bottleneck with 2 convolutions (nn/modules/block.py:187-207)."""

    def __init__(self, c1, c2, n=1, shortcut=False, g=1, e=0.5):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, self.c, shortcut, g, k=((3, 3), (3, 3)), e=1.0) for _ in range(n))

    def forward(self, x, out=None):
        self._check_mode(x)
        x = ops.as_act(x)
        n, _, h, w = x.shape
        c, nb = self.c, len(self.m)
        cat = ops.new_act(n, (2 + nb) * c, h, w, x.dtype, x.device)
        self.cv1(x, out=cat[:, :2 * c])
        for j, m in enumerate(self.m):
            m(cat[:, (1 + j) * c:(2 + j) * c], out=cat[:, (2 + j) * c:(3 + j) * c])
        return self.cv2(cat, out=out)

    forward_split = forward


class SPRModule(KernelModule):
    """Squeeze gate on [GAP(1) | GAP(2x2)] statistics (nn/modules/spr_module.py:8-31)."""

    def __init__(self, channels, reduction=4):
        super().__init__()
        self.avg_pool1 = nn.AdaptiveAvgPool2d(1)
        self.avg_pool2 = nn.AdaptiveAvgPool2d(2)
        self.fc1 = nn.Conv2d(channels * 5, channels // reduction, kernel_size=1, padding=0)
        self.relu = nn.ReLU(inplace=True)
        self.fc2 = nn.Conv2d(channels // reduction, channels, kernel_size=1, padding=0)
        self.sigmoid = nn.Sigmoid()

    def _pack(self, device):
        t = [self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias]
        return self._packed("mlp", torch.float32, device, t,
                            lambda: (f32(self.fc1.weight.flatten(1), device), f32(self.fc1.bias, device),
                                     f32(self.fc2.weight.flatten(1), device), f32(self.fc2.bias, device)))

    def forward(self, x):
        self._check_mode(x)
        x = ops.as_act(x)
        n, c, h, w = x.shape
        gate = ops.stats_gate(x, *self._pack(x.device), groups=1, softmax=False)   # statistics + MLP in one launch
        return gate.to(x.dtype).view(n, c, 1, 1)


class MSPA_C2f(KernelModule):
    """C2f with multi-scale (Res2Net-style) hierarchy and SPR channel attention
    (nn/modules/block.py:209-287)."""

    def __init__(self, inplanes, outplanes, n=1, shortcut=False, g=1, e=0.5, scale=4, stride=1, stype='normal'):
        super().__init__()
        self.nums = scale
        self.inwidth = inplanes // self.nums
        self.outwidth = outplanes // self.nums
        self.stride = stride
        assert stype in ['stage', 'normal'], 'One of these is suppported (stage or normal)'
        self.stype = stype
        self.convs = nn.ModuleList([])
        self.btnk_nums = n
        for i in range(self.nums):
            if self.stride == 1 and i != self.nums - 1:
                self.convs.append(Conv(self.inwidth, self.inwidth, 1, 1))
            else:
                self.convs.append(Conv(inplanes + self.outwidth * (n - 1), outplanes, 1, 1))
        self.bottleneck = nn.ModuleList(
            Bottleneck(self.inwidth, self.inwidth, shortcut, g, k=((3, 3), (3, 3)), e=1.0) for _ in range(n))
        self.attention = SPRModule(self.outwidth)
        self.softmax = nn.Softmax(dim=1)

    def _pack_front(self, dtype, device):
        """fp32 [g-1][ci][co] weights (BN folded, bf16-rounded values) + biases of the branch convs for
        mgdt_mspa_front, or None when the fused kernel does not take this block (fp32 mode, other widths)."""
        g, iw = self.nums, self.inwidth
        branch = list(self.convs)[:g - 1]
        acts = {act_name(c.act) for c in branch}
        if (dtype != torch.bfloat16 or not ops.FUSE_MSPA_FRONT or len(acts) != 1
                or not ops.lib().mgdt_mspa_front_supported(iw, g - 1)
                or any(c.conv.kernel_size != (1, 1) or c.conv.groups != 1 for c in branch)):
            return None
        tensors = []
        for c in branch:
            bn = getattr(c, "bn", None)
            tensors += [c.conv.weight] + ([c.conv.bias] if c.conv.bias is not None else [])
            if bn is not None:
                tensors += [bn.weight, bn.bias, bn.running_mean, bn.running_var]

        def build():
            ws, bs = zip(*(fold_conv_bn(c.conv, getattr(c, "bn", None)) for c in branch))
            w = torch.stack([t.flatten(1).to(device).to(torch.bfloat16).float().t().contiguous() for t in ws])
            return ops.mspa_front_pack(w.contiguous()), torch.stack([f32(b, device) for b in bs]).contiguous(), acts.copy().pop()

        return self._packed("front", dtype, device, tensors, build)

    def forward(self, x):
        self._check_mode(x)
        if self.stride != 1:
            raise NotImplementedError("MSPA_C2f: stride != 1 is not consistent in the reference either (block.py:233-237)")
        x = ops.as_act(x)
        n, c, h, w = x.shape
        g, iw, nb = self.nums, self.inwidth, self.btnk_nums
        if c != g * iw or (g - 1) * iw + nb * iw != self.convs[g - 1].conv.in_channels:
            raise ValueError("MSPA_C2f: channel arithmetic requires inplanes == outplanes, divisible by scale")
        cat = ops.new_act(n, (g - 1 + nb) * iw, h, w, x.dtype, x.device)
        front = self._pack_front(x.dtype, x.device) if g > 1 else None
        if front is not None:  # the whole pointwise branch chain + the last add in one launch
            sp = ops.mspa_front(x, front[0], front[1], iw, front[2], cat)
        else:
            sp = None
            for i in range(g - 1):  # sp = convs[i](sp + spx[i]); the add is fused into the conv's loader
                sp = self.convs[i](x[:, i * iw:(i + 1) * iw] if i == 0 else sp,
                                   pre_add=None if i == 0 else x[:, i * iw:(i + 1) * iw], out=cat[:, i * iw:(i + 1) * iw])
            last = x[:, (g - 1) * iw:]
            sp = ops.affine_act(sp, other=last) if g > 1 else last  # sp + spx[-1] feeds bottleneck + its shortcut
        for j, m in enumerate(self.bottleneck):
            sp = m(sp, out=cat[:, (g - 1 + j) * iw:(g + j) * iw])
        mlp = self.attention._pack(x.device)
        req = ops.gate_request(self.convs[g - 1].conv.out_channels, g, mlp[0].shape[0])
        feat = self.convs[g - 1](cat, stat=req)   # SPR statistics accumulated in the conv's epilogue when it can
        scale = ops.finish_gate(req, feat, *mlp, groups=g, softmax=True)
        return ops.affine_act(feat, a=scale)  # feats * softmax_g(gates), written back in group order


class SPPF(KernelModule):
    """Spatial pyramid pooling - fast (nn/modules/block.py:138-153)."""

    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * 4, c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)

    def forward(self, x):
        self._check_mode(x)
        x = ops.as_act(x)
        n, _, h, w = x.shape
        c_ = self.cv1.conv.out_channels
        cat = ops.new_act(n, 4 * c_, h, w, x.dtype, x.device)
        self.cv1(x, out=cat[:, :c_])
        ops.sppf_pool(cat[:, :c_], cat[:, c_:2 * c_], cat[:, 2 * c_:3 * c_], cat[:, 3 * c_:], self.m.kernel_size)
        return self.cv2(cat)


class SimFusion_4in(KernelModule):
    """Low-stage feature alignment of the GD neck (nn/modules/block.py:289-307): avg-pool the two
    larger maps, keep the third, bilinear-upsample the smallest, concat."""

    def __init__(self):
        super().__init__()
        self.avg_pool = nn.functional.adaptive_avg_pool2d

    def forward(self, x):
        self._check_mode(x)
        x_l, x_m, x_s, x_n = (ops.as_act(t) for t in x)
        n, _, h, w = x_s.shape
        parts = ((x_l, ops.RS_AVGPOOL), (x_m, ops.RS_AVGPOOL), (x_s, ops.RS_COPY), (x_n, ops.RS_BILINEAR))
        out = ops.new_act(n, sum(t.shape[1] for t, _ in parts), h, w, x_s.dtype, x_s.device)
        c0 = 0
        for t, mode in parts:
            ops.resample(t, h, w, mode, out=out[:, c0:c0 + t.shape[1]])
            c0 += t.shape[1]
        return out


class SimFusion_3in(KernelModule):
    """Lightweight adjacent-layer fusion (nn/modules/block.py:309-329); ReLU activations; branches
    whose Cin == Cout are nn.Identity (block.py:312-314)."""

    def __init__(self, in_channel_list, out_channels):
        super().__init__()
        self.cv1 = Conv(in_channel_list[0], out_channels, act=nn.ReLU()) if in_channel_list[0] != out_channels else nn.Identity()
        self.cv2 = Conv(in_channel_list[1], out_channels, act=nn.ReLU()) if in_channel_list[1] != out_channels else nn.Identity()
        self.cv3 = Conv(in_channel_list[2], out_channels, act=nn.ReLU()) if in_channel_list[2] != out_channels else nn.Identity()
        self.cv_fuse = Conv(out_channels * 3, out_channels, act=nn.ReLU())
        self.downsample = nn.functional.adaptive_avg_pool2d

    def forward(self, x):
        self._check_mode(x)
        x0, x1, x2 = (ops.as_act(t) for t in x)
        n, _, h, w = x1.shape
        co = self.cv_fuse.conv.out_channels
        cat = ops.new_act(n, 3 * co, h, w, x1.dtype, x1.device)
        for j, (t, mode, cv) in enumerate(((x0, ops.RS_AVGPOOL, self.cv1), (x1, ops.RS_COPY, self.cv2),
                                           (x2, ops.RS_BILINEAR, self.cv3))):
            dst = cat[:, j * co:(j + 1) * co]
            if isinstance(cv, nn.Identity):
                ops.resample(t, h, w, mode, out=dst)
            else:
                cv(t if mode == ops.RS_COPY else ops.resample(t, h, w, mode), out=dst)
        return self.cv_fuse(cat)


class LayerNorm(nn.Module):
    """Parameter container of the ConvNeXt LayerNorm (nn/modules/utils.py:145-169); the arithmetic
    is fused into mgdt_dwconv7_ln."""

    def __init__(self, normalized_shape, eps=1e-6, data_format="channels_last"):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps = eps
        self.data_format = data_format
        if self.data_format not in ["channels_last", "channels_first"]:
            raise NotImplementedError
        self.normalized_shape = (normalized_shape,)


class GRN(nn.Module):
    """Parameter container of global response normalisation (nn/modules/utils.py:171-182); folded
    into pwconv2 as a per-(n,c) input scale + bias."""

    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.zeros(1, 1, 1, dim))
        self.beta = nn.Parameter(torch.zeros(1, 1, 1, dim))


class ConvNeXtV2_Block(KernelModule):
    """ConvNeXt-V2 block (nn/modules/convnextv2.py:16-45) in four launches:
    dw7x7+LN | 1x1 expand + exact GELU | sum-of-squares per (n,c) | 1x1 project with the GRN scale
    on its input, GRN beta in its bias and the residual in its epilogue."""

    def __init__(self, dim, drop_path=0.):
        super().__init__()
        if drop_path > 0.:
            raise NotImplementedError("ConvNeXtV2_Block: drop_path > 0 is a training-only feature")
        self.dwconv = nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        self.grn = GRN(4 * dim)
        self.pwconv2 = nn.Linear(4 * dim, dim)
        self.drop_path = nn.Identity()

    def _pack(self, dtype, device):
        t = [self.dwconv.weight, self.dwconv.bias, self.norm.weight, self.norm.bias, self.pwconv1.weight,
             self.pwconv1.bias, self.grn.gamma, self.grn.beta, self.pwconv2.weight, self.pwconv2.bias]

        def build():
            c = self.dwconv.weight.shape[0]
            w2 = self.pwconv2.weight.detach().float()
            b2 = self.pwconv2.bias.detach().float() + w2 @ self.grn.beta.detach().float().reshape(-1)
            return dict(
                dw=self.dwconv.weight.detach().float().reshape(c, 49).t().contiguous().to(device=device, dtype=dtype),
                dwb=f32(self.dwconv.bias, device), lnw=f32(self.norm.weight, device), lnb=f32(self.norm.bias, device),
                w1=_pk1x1(self.pwconv1.weight.detach().float().reshape(4 * c, 1, 1, c), dtype, device),
                b1=f32(self.pwconv1.bias, device), gamma=f32(self.grn.gamma.reshape(-1), device),
                w2=_pk1x1(w2.reshape(c, 1, 1, 4 * c), dtype, device), b2=f32(b2, device))

        return self._packed("blk", dtype, device, t, build)

    def forward(self, x, out=None):
        self._check_mode(x)
        x = ops.as_act(x)
        p = self._pack(x.dtype, x.device)
        t = ops.dwconv7_ln(x, p["dw"], p["dwb"], p["lnw"], p["lnb"], self.norm.eps)
        req = ops.StatReq(0, True)
        hid = ops.conv2d(t, p["w1"], p["b1"], 1, act="gelu", stat=req)   # sum of squares accumulated in the epilogue
        scale = ops.finish_grn(req, hid, p["gamma"])
        return ops.conv2d(hid, p["w2"], p["b2"], 1, in_scale=scale, residual=x, out=out)


class IFM(KernelModule):
    """Information fusion module of the GD neck (nn/modules/block.py:331-342)."""

    def __init__(self, inc, ouc, embed_dim_p=96, fuse_block_num=3) -> None:
        super().__init__()
        self.conv = nn.Sequential(
            Conv(inc, embed_dim_p),
            *[ConvNeXtV2_Block(embed_dim_p) for _ in range(fuse_block_num)],
            Conv(embed_dim_p, sum(ouc)))

    def forward(self, x):
        self._check_mode(x)
        for m in self.conv:
            x = m(x)
        return x


class h_sigmoid(nn.Module):
    """relu6(x + 3) / 6 (nn/modules/block.py:344-350); fused into mgdt_inject."""

    def __init__(self, inplace=True):
        super().__init__()
        self.relu = nn.ReLU6(inplace=inplace)


class InjectionMultiSum_Auto_pool(KernelModule):
    """Information injection (nn/modules/block.py:352-399): local 1x1 embedding gated by the
    h-sigmoided, bilinearly upsampled global activation plus the upsampled global embedding."""

    def __init__(self, inp: int, oup: int, global_inp: list, flag: int) -> None:
        super().__init__()
        self.global_inp = global_inp
        self.flag = flag
        self.local_embedding = Conv(inp, oup, 1, act=False)
        self.global_embedding = Conv(global_inp[self.flag], oup, 1, act=False)
        self.global_act = Conv(global_inp[self.flag], oup, 1, act=False)
        self.act = h_sigmoid()

    def forward(self, x):
        self._check_mode(x)
        x_l, x_g = ops.as_act(x[0]), ops.as_act(x[1])
        c0 = sum(self.global_inp[:self.flag])
        info = x_g[:, c0:c0 + self.global_inp[self.flag]]  # x_g.split(global_inp, 1)[flag] (block.py:378)
        local = self.local_embedding(x_l)
        both = self._pack_global(info.dtype, info.device)
        pre_hsig = False
        if both is not None:   # global_act and global_embedding read the same input: one conv with stacked output channels
            w, b, oup = both
            n, _, h, wd = local.shape
            g = None
            if (info.dtype == torch.bfloat16 and oup % 32 == 0 and h == 2 * info.shape[2] and wd == 2 * info.shape[3]
                    and info.shape[2] > 1 and info.shape[3] > 1 and oup % 8 == 0):
                # h_sigmoid of the gate in the conv's epilogue (it precedes the interpolation, block.py:393): once per
                # low-resolution element instead of four times per output element inside mgdt_inject
                g = ops.conv2d(info, w, b, 1, act="hsigmoid", act_cols=oup)
                pre_hsig = g is not None
            if g is None:
                g = ops.conv2d(info, w, b, 1)
            gact, gfeat = g[:, :oup], g[:, oup:]
        else:
            gact = self.global_act(info)
            gfeat = self.global_embedding(info)
        return ops.inject(local, gact, gfeat, gact_is_hsig=pre_hsig)

    def _pack_global(self, dtype, device):
        """(PackedConv, bias, oup) of global_act and global_embedding stacked along the output channels (both are 1x1
        Conv+BN without activation on the same input, block.py:381-383), or None if they differ in shape."""
        ga, ge = self.global_act, self.global_embedding
        if (act_name(ga.act) is not None or act_name(ge.act) is not None or ga.conv.kernel_size != (1, 1)
                or ga.conv.weight.shape != ge.conv.weight.shape or ga.conv.groups != 1 or ge.conv.groups != 1):
            return None
        tensors = []
        for c in (ga, ge):
            bn = getattr(c, "bn", None)
            tensors += [c.conv.weight] + ([c.conv.bias] if c.conv.bias is not None else [])
            if bn is not None:
                tensors += [bn.weight, bn.bias, bn.running_mean, bn.running_var]

        def build():
            wa, ba = fold_conv_bn(ga.conv, getattr(ga, "bn", None))
            we, be = fold_conv_bn(ge.conv, getattr(ge, "bn", None))
            return ohwi(torch.cat([wa, we], 0), dtype, device, 1), f32(torch.cat([ba, be], 0), device), wa.shape[0]

        return self._packed("global2", dtype, device, tensors, build)


class _DeformWeights(nn.Module):
    """Parameter container standing where mmcv's ModulatedDeformConv2d stands (block.py:422-423)."""

    def __init__(self, cin, cout, k, stride=1, padding=1, bias=False):
        super().__init__()
        self.stride, self.padding = stride, padding
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        self.bias = nn.Parameter(torch.zeros(cout)) if bias else None


class DyDCNv2(KernelModule):
    """Modulated deformable 3x3 conv + GroupNorm(16) (nn/modules/block.py:401-432)."""

    def __init__(self, in_channels, out_channels, stride=1, norm_cfg=dict(type='GN', num_groups=16, requires_grad=True)):
        super().__init__()
        self.with_norm = norm_cfg is not None
        bias = not self.with_norm
        self.conv = _DeformWeights(in_channels, out_channels, 3, stride=stride, padding=1, bias=bias)
        if self.with_norm:
            if norm_cfg.get("type") != "GN":
                raise NotImplementedError("DyDCNv2: only GroupNorm is on the hot path")
            self.norm = nn.GroupNorm(norm_cfg["num_groups"], out_channels)

    def _pack(self, dtype, device):
        w = self.conv.weight
        t = [w] + ([self.norm.weight, self.norm.bias] if self.with_norm else [])

        def build():
            # [Cout][9][Cin] == OHWI of a 1x1 conv over 9*Cin virtual channels (the tensor-core DCN path)
            wp = _pk1x1(w.detach().float().permute(0, 2, 3, 1).reshape(w.shape[0], 1, 1, -1), dtype, device)
            if not self.with_norm:
                return wp, None, None
            return wp, f32(self.norm.weight, device), f32(self.norm.bias, device)

        return self._packed("dcn", dtype, device, t, build)

    def forward(self, x, offset, mask, mask_is_logit=False, act=None, out=None):
        """Reference signature forward(x, offset, mask); `mask_is_logit`/`act` let TOODHead pass channel
        slices of the raw offset conv and fuse the trailing F.relu (head.py:515-518,528)."""
        self._check_mode(x)
        if self.conv.stride != 1 or self.conv.bias is not None:
            raise NotImplementedError("DyDCNv2: stride 1 with a norm layer is the only configuration on the hot path")
        x = ops.as_act(x)
        offset, mask = ops.as_act(offset, x.dtype), ops.as_act(mask, x.dtype)
        wp, gnw, gnb = self._pack(x.dtype, x.device)
        req = ops.StatReq(1, True) if self.with_norm else None
        y = ops.dcn3x3(x, offset, mask, wp, self.conv.weight.shape[0], mask_is_logit, stat=req)
        if not self.with_norm:
            return y if act is None else ops.affine_act(y, act=act, out=out)
        a, b = ops.finish_gn(req, y, self.norm.num_groups, self.norm.eps, gnw, gnb)   # GroupNorm statistics from the conv's epilogue
        return ops.affine_act(y, a, b, act=act, out=out)
