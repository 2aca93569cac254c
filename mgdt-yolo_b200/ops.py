"""Tensor-level wrappers over the C ABI (torch is used for device memory and streams only).

Activations are torch CUDA tensors with logical shape (N, C, H, W) and NHWC physical layout
("channels_last", possibly a channel slice of a wider buffer).  `view(t)` validates that layout
and yields the raw pointer + channel stride the kernels take.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib
from ._lib import (ACT_GELU, ACT_HSIGMOID, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SILU, BF16, F32, RS_AVGPOOL,  # noqa: F401
                   RS_BILINEAR, RS_COPY, RS_NEAREST, ConvArgs, DecodeLevel, check, lib)

ACTS = {None: ACT_NONE, "none": ACT_NONE, "silu": ACT_SILU, "relu": ACT_RELU, "sigmoid": ACT_SIGMOID,
        "hsigmoid": ACT_HSIGMOID, "gelu": ACT_GELU}


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise TypeError(f"mgdt_b200 computes in float32 (validation) or bfloat16; got {dt}")


def require_cuda(t: torch.Tensor, what: str = "input"):
    if not t.is_cuda:
        raise RuntimeError(
            f"mgdt_b200: {what} is on {t.device}; this path runs only as sm_100a CUDA kernels -- there is no CPU "
            "fallback (move the model and tensors to a B200 device).")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def new_act(n, c, h, w, dtype, device) -> torch.Tensor:
    """(N,C,H,W)-shaped tensor with NHWC storage.  Channel counts >= 16 that are not a multiple of 8 (the heads' raw
    maps: 66 = 64 + nc, the 27 offset/mask channels) get a channel STRIDE rounded up to 8, so that pixels stay 16-byte
    aligned and the kernels keep their vector loads / stores (a 32->64 conv into a 66-channel map took 66 us with
    scalar stores, 3x the aligned time)."""
    cp = (c + 7) // 8 * 8 if c >= 16 else c
    t = torch.empty((n, h, w, cp), dtype=dtype, device=device)
    return (t if cp == c else t[..., :c]).permute(0, 3, 1, 2)


def _cs(t: torch.Tensor):
    """Channel stride of an NHWC-laid-out (N,C,H,W) tensor, or None if it is not one."""
    n, c, h, w = t.shape
    s0, s1, s2, s3 = t.stride()
    if w > 1:
        cs = s3
    elif h > 1:
        cs = s2
    elif n > 1:
        cs = s0
    else:
        cs = c
    ok = (cs >= c and (c == 1 or s1 == 1) and (w == 1 or s3 == cs) and (h == 1 or s2 == w * cs)
          and (n == 1 or s0 == h * w * cs))
    return cs if ok else None


def as_act(t: torch.Tensor, dtype: torch.dtype | None = None) -> torch.Tensor:
    """Bring any (N,C,H,W) tensor into NHWC storage (no copy if it already is)."""
    require_cuda(t)
    if t.dim() != 4:
        raise ValueError(f"expected a 4-D (N,C,H,W) tensor, got shape {tuple(t.shape)}")
    if dtype is None:
        dtype = torch.bfloat16 if t.dtype in (torch.bfloat16, torch.float16) else torch.float32
    if t.dtype != dtype:
        t = t.to(dtype)
    if _cs(t) is None:
        t = t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    return t


def view(t: torch.Tensor):
    """-> (ptr, N, C, H, W, channel_stride) of a conforming tensor."""
    cs = _cs(t)
    if cs is None:
        raise ValueError(f"tensor with shape {tuple(t.shape)} strides {t.stride()} is not NHWC-laid-out")
    n, c, h, w = t.shape
    return t.data_ptr(), n, c, h, w, cs


def _p(t):
    return None if t is None else t.data_ptr()


# Per-launch profiling (bench.py roofline): when PROFILE is a list every C-ABI call is bracketed by
# CUDA events on the current stream and appended as (name, meta, start, stop).  meta carries the
# ALGORITHMIC bytes / flops of the call (unique input elements + outputs + weights), not measured traffic.
PROFILE = None


def _nb(*ts):
    return sum(t.numel() * t.element_size() for t in ts if t is not None)


def _invoke(name, meta, *args):
    fn = getattr(lib(), name)
    if PROFILE is None:
        check(fn(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(fn(*args), name)
    e1.record()
    PROFILE.append((name, meta, e0, e1))


# ------------------------------------------------------------------------------------- conv
class PackedConv:
    """Weights of one convolution in the kernels' layouts: `.ohwi` (Cout,k,k,Cin) for the CUDA-core
    path and, when the tcgen05 path takes the shape (bf16 on a CUDA device), `.umma`, the K-major
    shared-memory image built by mgdt_conv_umma_pack for the stride it will be used with.  If the fp32
    master weights `w32` (same OHWI shape) are given and WEIGHT_F16 is on, the image is packed in fp16:
    the MMA then runs bf16 activations x fp16 weights, removing the weights' share of the rounding error."""

    def __init__(self, ohwi: torch.Tensor, stride: int = 1, w32: torch.Tensor | None = None):
        self.ohwi = ohwi
        self.umma = None
        self.f16 = False
        self.stem_u8 = None     # mgdt_stem_u8_pack image (layer 0 only, see stem_u8_pack)
        self.shape = ohwi.shape
        self.dtype = ohwi.dtype
        cout, k, k2, cin = ohwi.shape
        if ohwi.is_cuda and ohwi.dtype == torch.bfloat16 and k == k2 and USE_UMMA and lib().mgdt_has_umma():
            nbytes = lib().mgdt_conv_umma_packed_bytes(cin, cout, k, stride)
            if nbytes:
                self.umma = torch.empty((nbytes,), dtype=torch.uint8, device=ohwi.device)
                self.f16 = bool(WEIGHT_F16 and w32 is not None)
                src = w32.contiguous() if self.f16 else ohwi
                with torch.cuda.device(ohwi.device):
                    check(lib().mgdt_conv_umma_pack(src.data_ptr(), F32 if self.f16 else BF16, cin, cout, k, stride,
                                                    1 if self.f16 else 0, self.umma.data_ptr(), stream_ptr()),
                          "conv_umma_pack")
                self.stride = stride

    def data_ptr(self):
        return self.ohwi.data_ptr()


# tcgen05 path: pack weights as fp16 while activations stay bf16.  Measured on B200: tcgen05.mma kind::f16 with
# A = BF16 and B = F16 in the instruction descriptor raises 'illegal instruction' -- both operands must share one
# format -- so this stays off; the packer keeps the option for an all-fp16 mode.
WEIGHT_F16 = False
USE_UMMA = True  # tests flip this to compare the tcgen05 path with the CUDA-core path
PER_IMAGE_WEIGHTS = os.environ.get('MGDT_PER_IMAGE_W', '1') != '0'  # per-(n,c) input scales folded into per-image weights
FUSE_TOOD_SIBLINGS = os.environ.get('MGDT_FUSE_TOOD', '0') != '0'  # cls_decomp / reg_decomp / cls_prob_conv1 as one per-image-weight GEMM; measured SLOWER in an in-box A/B (19.5k -> 19.1k images/s: the DCN sampler then reads reg_feat as a slice of an 80-channel map, 108 -> 142 us), so off by default
STEM_MMA = os.environ.get('MGDT_STEM_MMA', '1') != '0'  # uint8 stem on warp-level MMAs (stem_mma.cu); 0: the tcgen05 stem
FUSE_MSPA_FRONT = True  # MSPA_C2f branch chain as one launch (bf16); tests flip this to compare with the unfused sequence


FUSE_STATS = os.environ.get('MGDT_FUSE_STATS', '1') != '0'  # per-(n,c) statistics in the tcgen05 conv's epilogue; tests flip this to compare with mgdt_chan_stats


class StatReq:
    """Request that a convolution accumulates per-(n, c) statistics of its output in its epilogue: q = 0 / 1 / 5 sum
    planes (none / total / total + adaptive 2x2 windows), sq = sum of squares.  After ops.conv2d returns, `fused` says
    whether the kernel took it (tcgen05 path) -- if not, the caller runs the stand-alone statistics pass."""

    def __init__(self, q, sq):
        self.q, self.sq, self.acc, self.fused = q, 1 if sq else 0, None, False


_STAT_ARENA = {}
_RETIRED = []   # outgrown arenas / ticket arrays, kept alive for CUDA graphs captured over them
STAT_COPIES = int(os.environ.get('MGDT_STAT_COPIES', '4'))  # replicas of the accumulators (copy = tile % copies): spreads the atomics over L2 lines


def _stat_arena(device, nelem):
    """Zero fp64 accumulators for the fused statistics, one array per (device, stream): the convolution adds into it,
    mgdt_stats_finish (the next launch on that stream) reads it and leaves it zero again.  Allocated outside any graph
    capture pool on first use (warm-up passes always precede capture)."""
    key = (device.index, stream_ptr())
    t = _STAT_ARENA.get(key)
    if t is None or t.numel() < nelem:
        if t is not None:
            _RETIRED.append(t)   # a captured graph may still add into the smaller arena: never free it
        t = torch.zeros((max(nelem, 1 << 16),), dtype=torch.float64, device=device)
        _STAT_ARENA[key] = t
    return t


def conv2d(x, w, bias, k, s=1, p=None, act=None, out=None, pre_add=None, in_scale=None, pix_scale=None,
           residual=None, in_relu=False, cout=None, impl=0, stat=None, scale_group_cols=None, act_cols=0):
    """y = act(conv((x + pre_add) * in_scale[n,c] * pix_scale[n,h,w] |> relu?, w) + bias) + residual.
    `w` is a PackedConv (or a plain OHWI (Cout, k, k, Cin) tensor) in x's dtype; `out` may be a channel
    slice of a concat buffer.  With `scale_group_cols`, in_scale is [G, N, Cin] and output columns
    [g*cols, (g+1)*cols) see the input scaled by in_scale[g] (columns past the last group: unscaled) -- sibling
    convs fused into one GEMM; only the per-image-weight tcgen05 path implements it (returns None otherwise).
    `act_cols` > 0: the activation applies to output channels < act_cols only (TMA-fed 1x1 kernel; returns None when
    another kernel would take the layer)."""
    xp, n, cin, h, wd, xcs = view(x)
    w_umma, w_f16 = None, False
    per_image = False
    if isinstance(w, PackedConv):
        if w.umma is not None and w.stride == s:
            w_umma, w_f16 = w.umma, w.f16
            per_image = (PER_IMAGE_WEIGHTS and in_scale is not None and pre_add is None and pix_scale is None and not in_relu
                         and impl != 1 and not w_f16 and x.dtype == torch.bfloat16)
        w = w.ohwi
    cout = cout if cout is not None else w.shape[0]
    p = k // 2 if p is None else p
    ho, wo = (h + 2 * p - k) // s + 1, (wd + 2 * p - k) // s + 1
    if out is None:
        out = new_act(n, cout, ho, wo, x.dtype, x.device)
    yp, yn, yc, yh, yw, ycs = view(out)
    if (yn, yc, yh, yw) != (n, cout, ho, wo):
        raise ValueError(f"conv2d: out shape {tuple(out.shape)} != {(n, cout, ho, wo)}")
    if w.dtype != x.dtype or out.dtype != x.dtype:
        raise TypeError("conv2d: x, w and out must share one dtype")
    a = ConvArgs()
    a.x, a.w, a.bias, a.y = xp, w.data_ptr(), _p(bias), yp
    a.N, a.H, a.W, a.Cin, a.Cout = n, h, wd, cin, cout
    a.kh = a.kw = k
    a.stride, a.pad = s, p
    a.x_cs, a.y_cs = xcs, ycs
    if pre_add is not None:
        ap, an, ac, ah, aw, acs = view(pre_add)
        if (an, ac, ah, aw) != (n, cin, h, wd) or pre_add.dtype != x.dtype:
            raise ValueError("conv2d: pre_add must match x")
        a.pre_add, a.add_cs = ap, acs
    ngroups = 1
    if in_scale is not None:
        if scale_group_cols is not None:
            ngroups = in_scale.shape[0]
            if not per_image or k != 1:
                return None
        if in_scale.dtype != torch.float32 or in_scale.numel() != ngroups * n * cin or not in_scale.is_contiguous():
            raise ValueError("conv2d: in_scale must be contiguous fp32 [N, Cin]")
        a.in_scale = in_scale.data_ptr()
    if pix_scale is not None:
        pp, pn, pc, ph, pw, pcs = view(pix_scale)
        if (pn, pc, ph, pw) != (n, 1, h, wd) or pix_scale.dtype != x.dtype:
            raise ValueError("conv2d: pix_scale must be (N,1,H,W) in x's dtype")
        a.pix_scale, a.ps_cs = pp, pcs
    if residual is not None:
        rp, rn, rc, rh, rw, rcs = view(residual)
        if (rn, rc, rh, rw) != (n, cout, ho, wo) or residual.dtype != x.dtype:
            raise ValueError("conv2d: residual must match the output")
        a.residual, a.res_cs = rp, rcs
    a.act = ACTS[act] if not isinstance(act, int) else act
    a.in_relu = 1 if in_relu else 0
    a.dtype = dtype_code(x.dtype)
    a.impl = impl
    a.w_umma = None if (w_umma is None or impl == 1) else w_umma.data_ptr()
    a.w_umma_f16 = 1 if w_f16 else 0
    if act_cols:
        a.act_cols = act_cols
        if lib().mgdt_conv2d_path(C.byref(a)) != 4:
            return None
    if per_image:
        # W (s_n o x) = (W diag(s_n)) x: if the tcgen05 path takes the layer with per-image weights (tiles cut per image,
        # weight slices carried through the ring), pack one scaled weight image per sample and run the transform-free
        # loader instead of scaling the activations inside the loader
        a.in_scale, a.w_per_image = None, 1
        if lib().mgdt_conv2d_path(C.byref(a)) in (2, 4, 5):
            pw = torch.empty((n * w_umma.numel(),), dtype=torch.uint8, device=x.device)
            _invoke("mgdt_conv_umma_pack_scaled_groups", dict(shape=f"pack_scaled {cin}->{cout} N{n}", bytes=pw.numel() + 4 * ngroups * n * cin,
                                                              flops=0.0, kernel="umma2_scale_packed_kernel"),
                    w_umma.data_ptr(), cin, cout, k, s, in_scale.data_ptr(), n, ngroups,
                    scale_group_cols if scale_group_cols is not None else (1 << 30), pw.data_ptr(), stream_ptr())
            a.w_umma = pw.data_ptr()
        elif scale_group_cols is not None:
            return None
        else:
            a.in_scale, a.w_per_image = in_scale.data_ptr(), 0
    es = x.element_size()
    meta = dict(shape=f"{cin}->{cout} k{k}s{s} {n}x{h}x{wd}", flops=2.0 * n * ho * wo * cout * cin * k * k,
                bytes=es * (n * h * wd * cin * (2 if pre_add is not None else 1)
                            + n * ho * wo * cout * (2 if residual is not None else 1) + cout * cin * k * k)
                + (n * h * wd * es if pix_scale is not None else 0))
    if stat is not None:
        stat.fused = False
        if FUSE_STATS and x.dtype == torch.bfloat16 and a.w_umma:
            stat.acc = _stat_arena(x.device, STAT_COPIES * n * (stat.q + stat.sq) * cout)
            a.stat_acc, a.stat_q, a.stat_sq, a.stat_copies = stat.acc.data_ptr(), stat.q, stat.sq, STAT_COPIES
            if lib().mgdt_conv2d_path(C.byref(a)) in (2, 4, 5):
                stat.fused = True
            else:
                a.stat_acc, a.stat_q, a.stat_sq, a.stat_copies = None, 0, 0, 0
    if PROFILE is not None:  # attribute the launch to the kernel the library will actually run
        meta["kernel"] = {3: "conv_pointwise_kernel", 2: "conv_umma2_kernel", 4: "conv1x1_tma_kernel", 5: "conv3x3_tma_kernel", 6: "conv3x3_warp_kernel"}.get(lib().mgdt_conv2d_path(C.byref(a)),
                                                                                  "conv_direct_kernel")
    _invoke("mgdt_conv2d", meta, C.byref(a), stream_ptr())
    return out


FUSE_TOOD_CLS = os.environ.get('MGDT_TOOD_CLS', '1') != '0'  # cls_prob_conv2 + sigmoid + cv3(cls_feat * cls_prob) as one launch


def tood_cls(prob, w2, b2, feat, w3, b3, out):
    """out = cv3(feat * sigmoid(conv3x3(prob))) (TOODHead, head.py:519-521, 528) in one launch; returns None when the
    fused kernel does not take the shape (the caller then runs the two convolutions)."""
    pp, n, c1, h, w, pcs = view(prob)
    fp, fn, c2, fh, fw, fcs = view(feat)
    op, on, nc, oh, ow, ocs = view(out)
    if (not FUSE_TOOD_CLS or prob.dtype != torch.bfloat16 or (fn, fh, fw) != (n, h, w) or (on, oh, ow) != (n, h, w)
            or tuple(w2.shape) != (1, 3, 3, c1) or tuple(w3.shape[:1] + w3.shape[-1:]) != (nc, c2) or w3.numel() != nc * c2
            or pp % 16 or fp % 16 or not lib().mgdt_tood_cls_supported(c1, c2, nc, pcs, fcs)):
        return None
    meta = dict(shape=f"tood_cls {c1}/{c2}->{nc} {n}x{h}x{w}", flops=2.0 * n * h * w * (9 * c1 + nc * c2),
                bytes=2 * n * h * w * (c1 + c2 + nc), kernel="tood_cls_kernel")
    _invoke("mgdt_tood_cls", meta, pp, pcs, w2.data_ptr(), _p(b2), fp, fcs, w3.data_ptr(), _p(b3), op, ocs, n, h, w, c1, c2, nc,
            dtype_code(prob.dtype), stream_ptr())
    return out


def mspa_front_pack(w32):
    """fp32 [nstage, iw(ci), iw(co)] CUDA weights -> the bf16 B-fragment image mgdt_mspa_front reads."""
    nstage, iw, _ = w32.shape
    nbytes = lib().mgdt_mspa_front_packed_bytes(iw, nstage)
    out = torch.empty((nbytes,), dtype=torch.uint8, device=w32.device)
    with torch.cuda.device(w32.device):
        check(lib().mgdt_mspa_front_pack(w32.contiguous().data_ptr(), nstage, iw, out.data_ptr(), stream_ptr()), "mspa_front_pack")
    return out


def mspa_front(x, w, bias, iw, act, ycat, ysp=None):
    """MSPA_C2f branch chain in one launch (bf16): x (N, (nstage+1)*iw, H, W); w = mspa_front_pack(weights);
    bias fp32 [nstage, iw].  Writes sp_i into ycat[:, i*iw:(i+1)*iw] and returns sp_in = sp_last + spx[-1]."""
    xp, n, c, h, wd, xcs = view(x)
    nstage = bias.shape[0]
    if c != (nstage + 1) * iw or x.dtype != torch.bfloat16:
        raise ValueError("mspa_front: x must be bf16 with (nstage+1)*iw channels")
    if ysp is None:
        ysp = new_act(n, iw, h, wd, x.dtype, x.device)
    yp, yn, yc, yh, yw, ycs = view(ycat)
    sp, sn, sc, sh, sw, scs = view(ysp)
    if (yn, yh, yw) != (n, h, wd) or yc < nstage * iw or (sn, sc, sh, sw) != (n, iw, h, wd):
        raise ValueError("mspa_front: bad output shapes")
    es = x.element_size()
    meta = dict(shape=f"mspa_front iw{iw}x{nstage} {n}x{h}x{wd}", flops=2.0 * n * h * wd * nstage * iw * iw,
                bytes=es * n * h * wd * (c + nstage * iw + iw) + w.numel(), kernel="mspa_front_kernel")
    _invoke("mgdt_mspa_front", meta, xp, xcs, w.data_ptr(), bias.data_ptr(), nstage, iw, ACTS[act], yp, ycs, sp, scs,
            n, h, wd, dtype_code(x.dtype), stream_ptr())
    return ysp


def stem_u8_pack(w32, cout):
    """B-fragment image of the stem's BN-folded fp32 weights (Cout, kp), k = (ky*3+kx)*3 + c, for mgdt_stem_u8; None if
    the warp-MMA stem does not take this width."""
    nbytes = lib().mgdt_stem_u8_packed_bytes(cout)
    if not nbytes or not STEM_MMA:
        return None
    w32 = w32.reshape(cout, -1).contiguous()
    out = torch.empty((nbytes,), dtype=torch.uint8, device=w32.device)
    with torch.cuda.device(w32.device):
        check(lib().mgdt_stem_u8_pack(w32.data_ptr(), w32.shape[1], cout, out.data_ptr(), stream_ptr()), "stem_u8_pack")
    return out


def stem_conv(src, w: "PackedConv", bias, cout, act, out=None):
    """Fused preprocess + 3x3/s2 stem conv from an NCHW uint8 (/255) or float32 image batch (bf16 out)."""
    require_cuda(src, "stem input")
    if not src.is_contiguous() or src.dtype not in (torch.uint8, torch.float32) or w.umma is None:
        raise ValueError("stem_conv: needs a contiguous uint8/float32 NCHW source and tcgen05-packed weights")
    n, c, h, wd = src.shape
    ho, wo = (h - 1) // 2 + 1, (wd - 1) // 2 + 1
    if out is None:
        out = new_act(n, cout, ho, wo, torch.bfloat16, src.device)
    yp, yn, yc, yh, yw, ycs = view(out)
    if (yn, yc, yh, yw) != (n, cout, ho, wo):
        raise ValueError("stem_conv: bad output shape")
    meta = dict(shape=f"stem {c}->{cout} k3s2 {n}x{h}x{wd}", flops=2.0 * n * ho * wo * cout * c * 9,
                bytes=_nb(src, out), kernel="conv_umma2_kernel")
    if (STEM_MMA and src.dtype == torch.uint8 and w.stem_u8 is not None and src.data_ptr() % 16 == 0
            and lib().mgdt_stem_u8_supported(c, h, wd, cout, ycs)):
        meta["kernel"] = "stem_mma_kernel"
        _invoke("mgdt_stem_u8", meta, src.data_ptr(), w.stem_u8.data_ptr(), _p(bias), yp, ycs, n, h, wd, cout, ACTS[act],
                stream_ptr())
        return out
    _invoke("mgdt_stem_conv", meta, src.data_ptr(), 1 if src.dtype == torch.uint8 else 0, w.umma.data_ptr(),
            1 if w.f16 else 0, _p(bias),
            yp, ycs, n, c, h, wd, cout, ACTS[act], BF16, stream_ptr())
    return out


def dwconv7_ln(x, w49c, bias, ln_w, ln_b, eps=1e-6, out=None):
    xp, n, c, h, w, xcs = view(x)
    if out is None:
        out = new_act(n, c, h, w, x.dtype, x.device)
    yp, *_, ycs = view(out)
    _invoke("mgdt_dwconv7_ln", dict(shape=f"dw7+ln C{c} {n}x{h}x{w}", bytes=_nb(x, out), flops=2.0 * 49 * x.numel()), xp, xcs, w49c.data_ptr(), bias.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), eps, yp,
                                ycs, n, h, w, c, dtype_code(x.dtype), stream_ptr())
    return out


def dcn3x3(x, offset, mask, w, cout, mask_is_logit, out=None, stat=None):
    xp, n, cin, h, wd, xcs = view(x)
    op, on, oc, oh, ow, ocs = view(offset)
    mp, mn, mc, mh, mw, mcs = view(mask)
    if (on, oc, oh, ow) != (n, 18, h, wd) or (mn, mc, mh, mw) != (n, 9, h, wd):
        raise ValueError("dcn3x3: offset must be (N,18,H,W) and mask (N,9,H,W)")
    if offset.dtype != x.dtype or mask.dtype != x.dtype:
        raise TypeError("dcn3x3: x, offset and mask must share one dtype")
    if out is None:
        out = new_act(n, cout, h, wd, x.dtype, x.device)
    yp, *_, ycs = view(out)
    umma_ptr = None if getattr(w, "umma", None) is None else w.umma.data_ptr()
    st = (None, 0, 0, 0)
    if stat is not None:
        stat.fused = False
        if FUSE_STATS and lib().mgdt_dcn3x3_path(xp, xcs, umma_ptr, n, h, wd, cin, cout, dtype_code(x.dtype)) == 2:
            stat.acc = _stat_arena(x.device, STAT_COPIES * n * (stat.q + stat.sq) * cout)
            st = (stat.acc.data_ptr(), stat.q, stat.sq, STAT_COPIES)
            stat.fused = True
    _invoke("mgdt_dcn3x3", dict(shape=f"dcn {cin}->{cout} {n}x{h}x{wd}", bytes=_nb(x, offset, mask, out), flops=2.0 * 9 * cin * cout * n * h * wd,
                                kernel="conv_umma2_kernel" if (getattr(w, "umma", None) is not None and x.dtype == torch.bfloat16) else "dcn3x3_kernel"), xp, xcs, op, ocs, mp, mcs, 1 if mask_is_logit else 0, w.data_ptr(),
            None if getattr(w, "umma", None) is None else w.umma.data_ptr(), 1 if getattr(w, "f16", False) else 0,
            yp, ycs, n, h, wd,
                            cin, cout, dtype_code(x.dtype), *st, stream_ptr())
    return out


# ------------------------------------------------------------------------------------- reductions
def chan_stats(x, quads=False, sumsq=False, fin=None):
    """-> (sum [N, 5|1, C] fp32, sumsq [N, C] fp32 | None).  `fin` (a _lib.StatsFin) makes the kernel's last block of
    every image also run the per-image consumer of the statistics (SPR gate / GRN scale / GroupNorm affine)."""
    xp, n, c, h, w, xcs = view(x)
    q = 5 if quads else 1
    s = torch.empty((n, q, c), dtype=torch.float32, device=x.device)
    ss = torch.empty((n, c), dtype=torch.float32, device=x.device) if sumsq else None
    nbytes = lib().mgdt_chan_stats_ws_bytes(n, h, w, c, 1 if quads else 0)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
    meta = dict(shape=f"stats C{c} {n}x{h}x{w} q{q}", bytes=_nb(x), flops=0.0)
    if fin is None:
        _invoke("mgdt_chan_stats", meta, xp, xcs, n, h, w, c, 1 if quads else 0, s.data_ptr(), _p(ss), ws.data_ptr(), nbytes,
                _stats_tickets(x.device, n).data_ptr(), dtype_code(x.dtype), stream_ptr())
    else:
        _invoke("mgdt_chan_stats_fin", meta, xp, xcs, n, h, w, c, 1 if quads else 0, s.data_ptr(), _p(ss), ws.data_ptr(), nbytes,
                _stats_tickets(x.device, n).data_ptr(), C.byref(fin), dtype_code(x.dtype), stream_ptr())
    return s, ss


def _gate_fits(c, groups, hidden, vec):
    ow = c // groups
    return groups * 5 * ow + groups * hidden + groups * ow <= 256 * (8 if vec else 1)


def stats_gate(x, fc1_w, fc1_b, fc2_w, fc2_b, groups=4, softmax=True):
    """SPRModule statistics + gate MLP (+ softmax over the groups) of x in ONE launch -> scale [N, C] fp32
    (== chan_stats(quads) followed by mspa_gate; falls back to the two launches when the scratch does not fit)."""
    from ._lib import StatsFin
    _, n, c, h, w, xcs = view(x)
    vec = c % 8 == 0 and xcs % 8 == 0 and x.data_ptr() % (8 * x.element_size()) == 0
    if not _gate_fits(c, groups, fc1_w.shape[0], vec):
        stats, _ = chan_stats(x, quads=True)
        return mspa_gate(stats, h, w, c, fc1_w, fc1_b, fc2_w, fc2_b, groups=groups, softmax=softmax)
    scale = torch.empty((n, c), dtype=torch.float32, device=x.device)
    fin = StatsFin(kind=1, p0=fc1_w.data_ptr(), p1=fc1_b.data_ptr(), p2=fc2_w.data_ptr(), p3=fc2_b.data_ptr(), i0=groups,
                   i1=1 if softmax else 0, i2=fc1_w.shape[0], f0=0.0, o0=scale.data_ptr(), o1=None)
    chan_stats(x, quads=True, fin=fin)
    return scale


def stats_grn(x, gamma):
    """sum of squares + GRN scale of x in one launch -> scale [N, C] fp32 (== chan_stats(sumsq) + grn_scale)."""
    from ._lib import StatsFin
    n, c = x.shape[0], x.shape[1]
    scale = torch.empty((n, c), dtype=torch.float32, device=x.device)
    fin = StatsFin(kind=2, p0=gamma.data_ptr(), p1=None, p2=None, p3=None, i0=0, i1=0, i2=0, f0=0.0, o0=scale.data_ptr(), o1=None)
    chan_stats(x, sumsq=True, fin=fin)
    return scale


def stats_gn(x, groups, eps, gamma, beta):
    """GroupNorm statistics + per-(n,c) affine (a, b) of x in one launch (== chan_stats(sumsq) + gn_affine)."""
    from ._lib import StatsFin
    n, c = x.shape[0], x.shape[1]
    a = torch.empty((n, c), dtype=torch.float32, device=x.device)
    b = torch.empty((n, c), dtype=torch.float32, device=x.device)
    fin = StatsFin(kind=3, p0=gamma.data_ptr(), p1=beta.data_ptr(), p2=None, p3=None, i0=groups, i1=0, i2=0, f0=float(eps),
                   o0=a.data_ptr(), o1=b.data_ptr())
    chan_stats(x, sumsq=True, fin=fin)
    return a, b


def _stats_finish(req, y, fin, want_sum=True):
    """mgdt_stats_finish on the accumulators a convolution filled for its output y -> (sum [N,q,C] | None, sumsq [N,C] | None)."""
    _, n, c, h, w, _ = view(y)
    s = torch.empty((n, req.q, c), dtype=torch.float32, device=y.device) if req.q else None
    ss = torch.empty((n, c), dtype=torch.float32, device=y.device) if req.sq else None
    _invoke("mgdt_stats_finish", dict(shape=f"stats_finish C{c} N{n} q{req.q}", bytes=8 * STAT_COPIES * n * (req.q + req.sq) * c, flops=0.0,
                                      kernel="stats_finish_kernel"),
            req.acc.data_ptr(), STAT_COPIES, n, h, w, c, req.q, req.sq, _p(s), _p(ss), None if fin is None else C.byref(fin), stream_ptr())
    return s, ss


def gate_request(c, groups, hidden):
    """StatReq for the SPR gate of a C-channel map if the fused finaliser's scratch fits, else None."""
    return StatReq(5, False) if _gate_fits(c, groups, hidden, True) else None


def finish_gate(req, y, fc1_w, fc1_b, fc2_w, fc2_b, groups=4, softmax=True):
    """SPR gate scale [N, C] of y: from the statistics its convolution accumulated (req.fused) or a stand-alone pass."""
    if req is None or not req.fused:
        return stats_gate(y, fc1_w, fc1_b, fc2_w, fc2_b, groups=groups, softmax=softmax)
    from ._lib import StatsFin
    n, c = y.shape[0], y.shape[1]
    scale = torch.empty((n, c), dtype=torch.float32, device=y.device)
    fin = StatsFin(kind=1, p0=fc1_w.data_ptr(), p1=fc1_b.data_ptr(), p2=fc2_w.data_ptr(), p3=fc2_b.data_ptr(), i0=groups,
                   i1=1 if softmax else 0, i2=fc1_w.shape[0], f0=0.0, o0=scale.data_ptr(), o1=None)
    _stats_finish(req, y, fin)
    return scale


def finish_grn(req, y, gamma):
    if req is None or not req.fused:
        return stats_grn(y, gamma)
    from ._lib import StatsFin
    n, c = y.shape[0], y.shape[1]
    scale = torch.empty((n, c), dtype=torch.float32, device=y.device)
    fin = StatsFin(kind=2, p0=gamma.data_ptr(), p1=None, p2=None, p3=None, i0=0, i1=0, i2=0, f0=0.0, o0=scale.data_ptr(), o1=None)
    _stats_finish(req, y, fin)
    return scale


def finish_gn(req, y, groups, eps, gamma, beta):
    if req is None or not req.fused:
        return stats_gn(y, groups, eps, gamma, beta)
    from ._lib import StatsFin
    n, c = y.shape[0], y.shape[1]
    a = torch.empty((n, c), dtype=torch.float32, device=y.device)
    b = torch.empty((n, c), dtype=torch.float32, device=y.device)
    fin = StatsFin(kind=3, p0=gamma.data_ptr(), p1=beta.data_ptr(), p2=None, p3=None, i0=groups, i1=0, i2=0, f0=float(eps),
                   o0=a.data_ptr(), o1=b.data_ptr())
    _stats_finish(req, y, fin)
    return a, b


_TICKETS = {}


def _stats_tickets(device, n):
    """Zero-initialised int32 tickets for the single-launch chan_stats (the kernel leaves them zero).  One array per
    (device, stream): calls on one stream are ordered, calls on different streams may overlap.  Allocated outside
    any graph capture pool on first use (warm-up passes always precede capture)."""
    key = (device.index, stream_ptr())
    t = _TICKETS.get(key)
    if t is None or t.numel() < n:
        if t is not None:
            _RETIRED.append(t)
        t = torch.zeros((max(n, 1024),), dtype=torch.int32, device=device)
        _TICKETS[key] = t
    return t


def mspa_gate(stats, h, w, c, fc1_w, fc1_b, fc2_w, fc2_b, groups=4, softmax=True):
    n = stats.shape[0]
    scale = torch.empty((n, c), dtype=torch.float32, device=stats.device)
    _invoke("mgdt_mspa_gate", dict(shape="gate", bytes=_nb(stats, scale), flops=0.0), stats.data_ptr(), n, h, w, c, groups, 1 if softmax else 0, fc1_w.data_ptr(),
                               fc1_b.data_ptr(), fc2_w.data_ptr(), fc2_b.data_ptr(), fc1_w.shape[0], scale.data_ptr(),
                               stream_ptr())
    return scale


def grn_scale(sumsq, gamma):
    n, c = sumsq.shape
    scale = torch.empty((n, c), dtype=torch.float32, device=sumsq.device)
    _invoke("mgdt_grn_scale", dict(shape="grn", bytes=_nb(sumsq, scale), flops=0.0), sumsq.data_ptr(), gamma.data_ptr(), n, c, scale.data_ptr(), stream_ptr())
    return scale


def gn_affine(s, ss, groups, hw, eps, gamma, beta):
    n, c = ss.shape
    a = torch.empty((n, c), dtype=torch.float32, device=ss.device)
    b = torch.empty((n, c), dtype=torch.float32, device=ss.device)
    _invoke("mgdt_gn_affine", dict(shape="gn", bytes=_nb(s, ss, a, b), flops=0.0), s.data_ptr(), ss.data_ptr(), n, c, groups, hw, eps, gamma.data_ptr(), beta.data_ptr(),
                               a.data_ptr(), b.data_ptr(), stream_ptr())
    return a, b


def td_attn(s, hw, la1_w, la1_b, la2_w, la2_b, stacked):
    """s [N,1,C] sums -> in_scale [ndec, N, C]; weights are stacked over the decompositions:
    la1_w [ndec, hidden, C], la1_b [ndec, hidden], la2_w [ndec, stacked, hidden], la2_b [ndec, stacked]."""
    n, c = s.shape[0], s.shape[-1]
    ndec, hidden = la1_w.shape[0], la1_w.shape[1]
    out = torch.empty((ndec, n, c), dtype=torch.float32, device=s.device)
    _invoke("mgdt_td_attn", dict(shape="td_attn", bytes=_nb(s, out), flops=0.0), s.data_ptr(), n, c, hw, hidden, stacked, ndec, la1_w.data_ptr(), la1_b.data_ptr(),
                             la2_w.data_ptr(), la2_b.data_ptr(), out.data_ptr(), stream_ptr())
    return out


# ------------------------------------------------------------------------------------- elementwise
def affine_act(x, a=None, b=None, act=None, other=None, out=None):
    xp, n, c, h, w, xcs = view(x)
    if out is None:
        out = new_act(n, c, h, w, x.dtype, x.device)
    yp, *_, ycs = view(out)
    op, ocs = None, 0
    if other is not None:
        op, on, oc, oh, ow, ocs = view(other)
        if (on, oc, oh, ow) != (n, c, h, w):
            raise ValueError("affine_act: other must match x")
    _invoke("mgdt_affine_act", dict(shape=f"affine C{c} {n}x{h}x{w}", bytes=_nb(x, out, other), flops=0.0), xp, xcs, _p(a), _p(b), op, ocs, ACTS[act], yp, ycs, n, h, w, c, dtype_code(x.dtype),
                                stream_ptr())
    return out


def resample(x, ho, wo, mode, out=None):
    xp, n, c, h, w, xcs = view(x)
    if out is None:
        out = new_act(n, c, ho, wo, x.dtype, x.device)
    yp, yn, yc, yh, yw, ycs = view(out)
    if (yn, yc, yh, yw) != (n, c, ho, wo):
        raise ValueError(f"resample: out shape {tuple(out.shape)} != {(n, c, ho, wo)}")
    _invoke("mgdt_resample", dict(shape=f"resample m{mode} C{c} {h}x{w}->{ho}x{wo}", bytes=_nb(x, out), flops=0.0), xp, xcs, h, w, yp, ycs, ho, wo, n, c, mode, dtype_code(x.dtype), stream_ptr())
    return out


def sppf_pool(x, y1, y2, y3, k=5):
    xp, n, c, h, w, xcs = view(x)
    p1, *_, ycs = view(y1)
    p2, *_, ycs2 = view(y2)
    p3, *_, ycs3 = view(y3)
    assert ycs == ycs2 == ycs3
    _invoke("mgdt_sppf_pool", dict(shape=f"sppf C{c} {n}x{h}x{w}", bytes=_nb(x) * 4, flops=0.0), xp, xcs, p1, p2, p3, ycs, n, h, w, c, k, dtype_code(x.dtype), stream_ptr())


def inject(local, gact, gfeat, out=None, gact_is_hsig=False):
    lp, n, c, h, w, lcs = view(local)
    ap, an, ac, hg, wg, acs = view(gact)
    fp, fn, fc, fh, fw, fcs = view(gfeat)
    if (an, ac) != (n, c) or (fn, fc, fh, fw) != (an, ac, hg, wg):
        raise ValueError("inject: shape mismatch")
    if out is None:
        out = new_act(n, c, h, w, local.dtype, local.device)
    yp, *_, ycs = view(out)
    _invoke("mgdt_inject2", dict(shape=f"inject C{c} {n}x{h}x{w}", bytes=_nb(local, gact, gfeat, out), flops=0.0, kernel="inject_kernel"), lp, lcs, ap, acs, fp, fcs, yp, ycs,
            n, h, w, hg, wg, c, 1 if gact_is_hsig else 0, dtype_code(local.dtype), stream_ptr())
    return out


def preprocess(src: torch.Tensor, dtype: torch.dtype, out=None):
    """uint8 or float32 NCHW-contiguous (N,C,H,W) -> NHWC `dtype`; uint8 is divided by 255."""
    require_cuda(src)
    if not src.is_contiguous() or src.dtype not in (torch.uint8, torch.float32):
        raise ValueError("preprocess: source must be a contiguous uint8/float32 NCHW tensor")
    n, c, h, w = src.shape
    if out is None:
        out = new_act(n, c, h, w, dtype, src.device)
    yp, *_, ycs = view(out)
    _invoke("mgdt_preprocess", dict(shape=f"preprocess {n}x{c}x{h}x{w}", bytes=_nb(src, out), flops=0.0), src.data_ptr(), 1 if src.dtype == torch.uint8 else 0, yp, ycs, n, c, h, w,
                                dtype_code(dtype), stream_ptr())
    return out


# ------------------------------------------------------------------------------------- decode / NMS
def decode(raws, strides, reg_max, nc, out=None, dist_only=False):
    n = raws[0].shape[0]
    levels = (DecodeLevel * len(raws))()
    total = 0
    for i, (r, s) in enumerate(zip(raws, strides)):
        rp, rn, rc, rh, rw, rcs = view(r)
        if rc != 4 * reg_max + nc:
            raise ValueError("decode: raw map has the wrong channel count")
        levels[i].raw, levels[i].H, levels[i].W, levels[i].cs, levels[i].stride = rp, rh, rw, rcs, float(s)
        total += rh * rw
    if out is None:
        out = torch.empty((n, 4 if dist_only else 4 + nc, total), dtype=torch.float32, device=raws[0].device)
    _invoke("mgdt_decode", dict(shape=f"decode A{total} nc{nc}", bytes=_nb(*raws) + _nb(out), flops=0.0), levels, len(raws), n, reg_max, nc, 1 if dist_only else 0, out.data_ptr(),
                            dtype_code(raws[0].dtype),
                            stream_ptr())
    return out


def nms_packed(pred, conf_thres, iou_thres, multi_label=False, agnostic=False, max_det=300, max_nms=30000,
               max_wh=7680.0, classes=None, out=None, counts=None, ws=None):
    """-> (out [N, max_det, 6] fp32, counts [N] int32), all on device, no synchronisation."""
    require_cuda(pred, "prediction")
    if pred.dtype != torch.float32 or not pred.is_contiguous():
        pred = pred.float().contiguous()
    n, ch, a = pred.shape
    nc = ch - 4
    nbytes = lib().mgdt_nms_ws_bytes(n, nc, a, 1 if multi_label else 0, max_nms)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=pred.device)
    if out is None:
        out = torch.empty((n, max_det, 6), dtype=torch.float32, device=pred.device)
    if counts is None:
        counts = torch.empty((n,), dtype=torch.int32, device=pred.device)
    cls_t = None
    if isinstance(classes, torch.Tensor):
        cls_t = classes  # int32, on the device already (graph-capturable)
    elif classes is not None:
        cls_t = torch.as_tensor(list(classes), dtype=torch.int32, device=pred.device)
    _invoke("mgdt_nms", dict(shape=f"nms N{n} nc{nc} A{a}", bytes=_nb(pred, out), flops=0.0), pred.data_ptr(), n, nc, a, float(conf_thres), float(iou_thres), 1 if multi_label else 0,
                         1 if agnostic else 0, max_det, max_nms, float(max_wh), _p(cls_t),
                         0 if cls_t is None else cls_t.numel(), out.data_ptr(), counts.data_ptr(), ws.data_ptr(),
                         ws.numel(), stream_ptr())
    return out, counts


# ------------------------------------------------------------------------------------- pre / post-processing
def letterbox_u8(src, dst, new_hw, top_left, swap_rb=True, pad_value=114, out_hwc=False):
    """One HWC uint8 image (CUDA, 3 channels, contiguous rows) -> `dst`: a (3, H, W) slot of an NCHW uint8 batch
    (or an (H, W, 3) image when out_hwc), resized to new_hw at offset top_left, the rest filled with pad_value."""
    require_cuda(src, "image")
    if src.dtype != torch.uint8 or src.dim() != 3 or src.shape[2] != 3 or src.stride(2) != 1 or src.stride(1) != 3:
        raise ValueError("letterbox_u8: expects an (h, w, 3) uint8 image with contiguous pixels")
    if dst.dtype != torch.uint8 or not dst.is_contiguous():
        raise ValueError("letterbox_u8: dst must be a contiguous uint8 tensor")
    h0, w0 = int(src.shape[0]), int(src.shape[1])
    H, W = (int(dst.shape[0]), int(dst.shape[1])) if out_hwc else (int(dst.shape[1]), int(dst.shape[2]))
    _invoke("mgdt_letterbox_u8", dict(shape=f"letterbox {h0}x{w0}->{H}x{W}", bytes=_nb(src, dst), flops=0.0), src.data_ptr(),
            h0, w0, int(src.stride(0)), dst.data_ptr(), H, W, int(new_hw[0]), int(new_hw[1]), int(top_left[0]),
            int(top_left[1]), 1 if swap_rb else 0, int(pad_value), 1 if out_hwc else 0, stream_ptr())
    return dst


def scale_boxes_packed(dets, counts, params):
    """In place: dets (N, max_rows, >=4) fp32 xyxy..., counts (N,) int32 or None, params (N, 5) fp32
    [gain, pad_w, pad_h, h0, w0] -> boxes mapped back to the original images and clipped."""
    require_cuda(dets, "detections")
    if dets.dtype != torch.float32 or not dets.is_contiguous() or dets.dim() != 3:
        raise ValueError("scale_boxes_packed: dets must be a contiguous fp32 (N, rows, >=4) tensor")
    n, rows, row = dets.shape
    if rows == 0:
        return dets
    _invoke("mgdt_scale_boxes", dict(shape=f"scale_boxes N{n}", bytes=_nb(dets), flops=0.0), dets.data_ptr(), row, _p(counts),
            n, rows, params.data_ptr(), stream_ptr())
    return dets


def box_convert(boxes, mode, w=1.0, h=1.0):
    """xyxy columns of an fp32 CUDA (n, >=4) tensor -> (n, 4): mode bit 0 xyxy->xywh, bit 1 normalise by (w, h)."""
    require_cuda(boxes, "boxes")
    if boxes.dtype != torch.float32 or boxes.dim() != 2 or boxes.shape[1] < 4 or boxes.stride(1) != 1:
        raise ValueError("box_convert: expects an fp32 (n, >=4) tensor with contiguous rows")
    n = boxes.shape[0]
    out = torch.empty((n, 4), dtype=torch.float32, device=boxes.device)
    _invoke("mgdt_box_convert", dict(shape=f"box_convert n{n}", bytes=_nb(out) * 2, flops=0.0), boxes.data_ptr(),
            int(boxes.stride(0)) if n > 1 else boxes.shape[1], n, mode, float(w), float(h), out.data_ptr(), stream_ptr())
    return out


def match_batch(dets, det_counts, labels, lab_counts, iouv):
    """dets (N, max_det, >=6) fp32, labels (N, max_lab, 5) fp32 (cls, xyxy), counts int32 (N,) or None, iouv fp32 (niou,)
    -> correct (N, max_det, niou) bool, all on device, no synchronisation."""
    require_cuda(dets, "detections")
    if dets.dtype != torch.float32 or not dets.is_contiguous() or dets.dim() != 3 or dets.shape[2] < 6:
        raise ValueError("match_batch: dets must be a contiguous fp32 (N, max_det, >=6) tensor")
    if labels.dtype != torch.float32 or not labels.is_contiguous() or labels.dim() != 3 or labels.shape[2] != 5:
        raise ValueError("match_batch: labels must be a contiguous fp32 (N, max_lab, 5) tensor")
    n, max_det, row = dets.shape
    max_lab, niou = labels.shape[1], iouv.numel()
    correct = torch.zeros((n, max_det, niou), dtype=torch.uint8, device=dets.device)
    _invoke("mgdt_match_batch", dict(shape=f"match N{n} D{max_det} L{max_lab}", bytes=_nb(dets, labels, correct), flops=0.0),
            dets.data_ptr(), row, _p(det_counts), max_det, labels.data_ptr(), _p(lab_counts), max_lab, iouv.data_ptr(), niou,
            correct.data_ptr(), n, stream_ptr())
    return correct.bool()
