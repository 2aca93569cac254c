"""GPU: the tcgen05/TMEM convolution path against the CUDA-core path and the oracle.

Both paths see identical bf16 inputs/weights and accumulate in fp32, so they must agree to bf16
output rounding (1 ulp = 2^-8 relative) -- far tighter than the 1e-2 budget against the fp32 oracle.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (Cin, Cout, k, s, N, H, W)
SHAPES = [
    (32, 32, 1, 1, 2, 20, 24), (8, 8, 1, 1, 3, 17, 19), (16, 16, 1, 1, 2, 40, 40), (96, 384, 1, 1, 2, 12, 10),
    (384, 96, 1, 1, 2, 12, 10), (480, 96, 1, 1, 1, 9, 11), (64, 256, 1, 1, 2, 16, 16), (256, 64, 1, 1, 2, 16, 16),
    (192, 64, 1, 1, 1, 8, 9), (32, 2, 1, 1, 2, 10, 10), (32, 64, 1, 1, 1, 80, 80), (512, 256, 1, 1, 2, 6, 7),
    (8, 8, 3, 1, 2, 21, 19), (16, 16, 3, 1, 2, 16, 24), (32, 32, 3, 1, 2, 20, 20), (64, 64, 3, 1, 2, 9, 11),
    (64, 32, 3, 1, 1, 24, 20), (64, 27, 3, 1, 2, 12, 16), (16, 1, 3, 1, 2, 14, 10), (32, 32, 3, 1, 1, 80, 80),
    (16, 32, 3, 2, 2, 32, 40), (32, 64, 3, 2, 2, 20, 24), (64, 128, 3, 2, 2, 16, 12), (16, 32, 3, 2, 1, 21, 23),
    (8, 16, 3, 2, 2, 18, 14), (128, 256, 3, 2, 1, 8, 8),
    # large enough for 256/512-row tiles: Cout <= 16 runs the PAIRED 16-column epilogue units (two row blocks per unit)
    (8, 8, 3, 1, 5, 160, 160), (16, 16, 3, 1, 8, 120, 130), (64, 16, 1, 1, 8, 160, 100), (16, 8, 3, 1, 3, 200, 212),
    # column splits whose last unit has 16 columns (Cout 576 -> 4 x 144): that unit must not be stored as a 32-channel TMA box
    (64, 576, 1, 1, 2, 12, 10), (96, 48, 1, 1, 3, 20, 24), (32, 27, 1, 1, 2, 30, 30),
    # the TMA-fed 1x1 kernel at sizes where every persistent CTA walks several 128-row tiles, the ring wraps, two CTAs
    # share an SM and the last tile is partial; K blocks of 64 / 32 / 16 channels (SWIZZLE_128B / 64B / 32B)
    (64, 256, 1, 1, 8, 80, 80), (96, 384, 1, 1, 8, 40, 40), (32, 32, 1, 1, 4, 160, 160), (160, 128, 1, 1, 4, 40, 40),
    (80, 64, 1, 1, 4, 80, 80), (384, 96, 1, 1, 8, 40, 40), (512, 256, 1, 1, 8, 20, 20), (32, 512, 1, 1, 4, 40, 40),
    (48, 40, 1, 1, 3, 33, 37), (16, 24, 1, 1, 5, 50, 61),
]


def _mk(cin, cout, k, n, h, w, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    return x, wt, b


@pytest.mark.parametrize("shape", SHAPES, ids=[f"{s[0]}to{s[1]}k{s[2]}s{s[3]}_{s[4]}x{s[5]}x{s[6]}" for s in SHAPES])
def test_umma_vs_direct_and_oracle(shape):
    from mgdt_yolo_b200 import ops
    from mgdt_yolo_b200._lib import lib
    if not lib().mgdt_has_umma():
        pytest.skip("library built without the tcgen05 path")
    cin, cout, k, s, n, h, w = shape
    x, wt, b = _mk(cin, cout, k, n, h, w)
    xb = ops.as_act(x.cuda().to(torch.bfloat16))
    pw = ops.PackedConv(wt.cuda().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16), s)
    if pw.umma is None:
        pytest.skip("shape not taken by the tcgen05 path")
    bias = b.cuda()
    try:
        y_umma = ops.conv2d(xb, pw, bias, k, s, act="silu", impl=2)
    except RuntimeError as e:
        if "does not support this shape" in str(e):
            pytest.skip("operands do not fit the whole-K-resident tile (falls back to the CUDA-core path)")
        raise
    y_dir = ops.conv2d(xb, pw, bias, k, s, act="silu", impl=1)
    torch.cuda.synchronize()
    ref = F.silu(F.conv2d(xb.float(), pw.ohwi.float().permute(0, 3, 1, 2), bias, stride=s, padding=k // 2))
    scale = float(ref.abs().max())
    e_ud = float((y_umma.float() - y_dir.float()).abs().max()) / scale
    e_uo = float((y_umma.float() - ref).abs().max()) / scale
    assert y_umma.shape == ref.shape
    assert e_ud <= 2 ** -7, f"tcgen05 vs CUDA-core: {e_ud:.3e}"
    assert e_uo <= 2 ** -7, f"tcgen05 vs fp32 reference on identical bf16 operands: {e_uo:.3e}"


def test_umma_fused_options():
    """pre_add, in_scale, pix_scale, in_relu, residual, channel-slice input/output."""
    from mgdt_yolo_b200 import ops
    from mgdt_yolo_b200._lib import lib
    if not lib().mgdt_has_umma():
        pytest.skip("library built without the tcgen05 path")
    g = torch.Generator().manual_seed(5)
    n, h, w, cin, cout = 2, 18, 22, 32, 48
    xbuf = ops.as_act(torch.randn(n, 80, h, w, generator=g).cuda().to(torch.bfloat16))
    x = xbuf[:, 16:48]
    add = ops.as_act(torch.randn(n, cin, h, w, generator=g).cuda().to(torch.bfloat16))
    res = ops.as_act(torch.randn(n, cout, h, w, generator=g).cuda().to(torch.bfloat16))
    pix = ops.as_act(torch.rand(n, 1, h, w, generator=g).cuda().to(torch.bfloat16))
    insc = (torch.rand(n, cin, generator=g) + 0.5).cuda().contiguous()
    bias = torch.randn(cout, generator=g).cuda()
    for k in (1, 3):
        wt = (torch.randn(cout, k, k, cin, generator=g) * (2.0 / (cin * k * k)) ** 0.5).cuda().to(torch.bfloat16)
        pw = ops.PackedConv(wt, 1)
        outs = []
        for impl in (2, 1):
            obuf = ops.new_act(n, 96, h, w, torch.bfloat16, "cuda")
            obuf.zero_()
            ops.conv2d(x, pw, bias, k, 1, act="relu", out=obuf[:, 32:80], pre_add=add, in_scale=insc, pix_scale=pix,
                       residual=res, in_relu=True, impl=impl)
            outs.append(obuf.float())
        torch.cuda.synchronize()
        assert float(outs[0][:, :32].abs().max()) == 0 and float(outs[0][:, 80:].abs().max()) == 0
        a = torch.relu((x.float() + add.float()) * insc.view(n, cin, 1, 1) * pix.float()).to(torch.bfloat16).float()
        ref = F.relu(F.conv2d(a, wt.float().permute(0, 3, 1, 2), bias, padding=k // 2)) + res.float()
        scale = float(ref.abs().max())
        assert float((outs[0] - outs[1]).abs().max()) / scale <= 2 ** -6
        assert float((outs[0][:, 32:80] - ref).abs().max()) / scale <= 2 ** -6


@pytest.mark.parametrize("cin,cout", [(8, 8), (16, 16), (8, 16), (16, 8)])
def test_pointwise_narrow_conv(cin, cout):
    """The HBM-bound SIMT kernel for narrow 1x1 layers (impl=0 dispatch) against the CUDA-core reference path,
    with residual, pre_add and channel-slice operands."""
    from mgdt_yolo_b200 import ops
    from mgdt_yolo_b200._lib import lib
    import ctypes as C
    g = torch.Generator().manual_seed(cin * 100 + cout)
    n, h, w = 3, 19, 23
    xbuf = ops.as_act(torch.randn(n, cin + 16, h, w, generator=g).cuda().to(torch.bfloat16))
    x = xbuf[:, 8:8 + cin]
    add = ops.as_act(torch.randn(n, cin, h, w, generator=g).cuda().to(torch.bfloat16))
    res = ops.as_act(torch.randn(n, cout, h, w, generator=g).cuda().to(torch.bfloat16))
    wt = (torch.randn(cout, 1, 1, cin, generator=g) * (2.0 / cin) ** 0.5).cuda().to(torch.bfloat16)
    pw = ops.PackedConv(wt, 1)
    bias = torch.randn(cout, generator=g).cuda()
    ybuf = ops.as_act(torch.zeros(n, cout + 8, h, w).cuda().to(torch.bfloat16))
    for kw in (dict(), dict(residual=res), dict(pre_add=add), dict(residual=res, pre_add=add)):
        for act in ("silu", None, "relu"):
            y0 = ops.conv2d(x, pw, bias, 1, 1, act=act, impl=0, out=ybuf[:, 8:], **kw)
            y1 = ops.conv2d(x, pw, bias, 1, 1, act=act, impl=1, **kw)
            torch.cuda.synchronize()
            scale = float(y1.float().abs().max())
            err = float((y0.float() - y1.float()).abs().max()) / scale
            assert err <= 2 ** -7, f"{kw.keys()} {act}: {err:.3e}"
    assert float(ybuf[:, :8].float().abs().max()) == 0.0   # the slice write stayed inside its channels


@pytest.mark.parametrize("iw,n", [(8, 1), (16, 2), (32, 2), (64, 1)])
def test_mspa_front_fused_vs_unfused(iw, n):
    """mgdt_mspa_front (the MSPA_C2f branch chain in one launch) against the unfused sequence of pointwise convs with
    pre_add + affine_act, through the module, on a ragged shape and as part of the whole block."""
    from mgdt_yolo_b200 import ops
    from mgdt_yolo_b200.modules.block import MSPA_C2f
    from mgdt_yolo_b200.synth import synth_state_dict
    torch.manual_seed(iw)
    m = MSPA_C2f(4 * iw, 4 * iw, n, True)
    m.load_state_dict(synth_state_dict(m.state_dict(), seed=iw))
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(iw + 1)
    x = ops.as_act(torch.randn(3, 4 * iw, 13, 17, generator=g).cuda().to(torch.bfloat16))
    outs, launches = {}, {}
    for rnd in range(2):   # round 0 packs the weights (its kernels count as launches too); round 1 is the one compared
        for fused in (True, False):
            ops.FUSE_MSPA_FRONT = fused
            try:
                before = ops.lib().mgdt_launch_count()
                with torch.no_grad():
                    outs[fused] = m(x).float()
                torch.cuda.synchronize()
                launches[fused] = ops.lib().mgdt_launch_count() - before
            finally:
                ops.FUSE_MSPA_FRONT = True
    assert launches[True] == launches[False] - 3, launches   # 3 convs + 1 add -> 1 launch
    scale = float(outs[False].abs().max())
    err = float((outs[True] - outs[False]).abs().max()) / scale
    assert err <= 2 ** -7, f"iw={iw}: {err:.3e}"


def _fused_epilogue_stats_case(kind, shape):
    """Per-(n,c) statistics accumulated in the tcgen05 conv's epilogue (fp64 atomics + mgdt_stats_finish) against the
    stand-alone mgdt_chan_stats pass, through the modules that use them: SPR gate (2x2 adaptive windows, odd sizes ->
    overlapping windows, images that straddle 32-row groups), GRN (sum of squares), GroupNorm (sum + sum of squares).
    Run three times: the accumulators must be left zero for the next use and the result must be reproducible."""
    from mgdt_yolo_b200 import ops
    from mgdt_yolo_b200.modules.block import ConvNeXtV2_Block, MSPA_C2f
    from mgdt_yolo_b200.modules.head import Conv_GN
    from mgdt_yolo_b200.synth import synth_state_dict
    n, c, h, w = shape
    torch.manual_seed(c + h)
    m = {"mspa": lambda: MSPA_C2f(c, c, 1, True), "convnext": lambda: ConvNeXtV2_Block(c),
         "conv_gn": lambda: Conv_GN(c, c, 3)}[kind]()
    m.load_state_dict(synth_state_dict(m.state_dict(), seed=h))
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(h * w)
    x = ops.as_act(torch.randn(n, c, h, w, generator=g).cuda().to(torch.bfloat16))
    outs, launches = {}, {}
    for rnd in range(3):
        for fused in (True, False):
            ops.FUSE_STATS = fused
            try:
                before = ops.lib().mgdt_launch_count()
                with torch.no_grad():
                    y = m(x).float()
                torch.cuda.synchronize()
                launches[fused] = ops.lib().mgdt_launch_count() - before
            finally:
                ops.FUSE_STATS = True
            if rnd > 0 and fused:
                assert torch.equal(y, outs[True]), "fused statistics are not reproducible / accumulators not reset"
            outs[fused] = y
    assert launches[True] == launches[False], launches   # chan_stats launch -> stats_finish launch
    for t in ops._STAT_ARENA.values():
        assert float(t.abs().max()) == 0.0
    scale = float(outs[False].abs().max())
    err = float((outs[True] - outs[False]).abs().max()) / scale
    assert err <= 2 ** -7, f"{kind} {shape}: {err:.3e}"


@pytest.mark.parametrize("kind,shape", [("mspa", (3, 128, 13, 17)), ("mspa", (2, 128, 20, 20)), ("mspa", (5, 256, 8, 8)),
                                        ("convnext", (3, 96, 13, 17)), ("convnext", (2, 96, 40, 40)),
                                        ("conv_gn", (3, 64, 13, 17)), ("conv_gn", (2, 64, 80, 80))])
def test_fused_epilogue_stats_vs_chan_stats(kind, shape):
    _fused_epilogue_stats_case(kind, shape)


@pytest.mark.parametrize("kind,shape", [("mspa", (2, 32, 20, 28)), ("mspa", (3, 128, 37, 41)), ("convnext", (2, 96, 40, 40)),
                                        ("convnext", (5, 96, 23, 17))])
def test_fused_stats_in_tma_kernel(kind, shape):
    """The same check with the 1x1 statistics layers routed to the TMA-fed kernel (option conv_tma_stats; off by default
    because it measured slower): per-warp shared-memory accumulators flushed when the image changes."""
    from mgdt_yolo_b200._lib import lib
    lib().mgdt_set_option(b"conv_tma_stats", 1)
    try:
        _fused_epilogue_stats_case(kind, shape)
    finally:
        lib().mgdt_set_option(b"conv_tma_stats", 0)


@pytest.mark.parametrize("cin,cout,k", [(16, 16, 3), (8, 8, 3), (32, 16, 1)])
def test_umma_paired_units_with_residual(cin, cout, k):
    """Paired 16-column epilogue units (Cout <= 16, multi-row-block tiles) with a residual and a channel-slice output:
    the second half of a unit reads its residual / writes its rows from the NEXT row block."""
    from mgdt_yolo_b200 import ops
    g = torch.Generator().manual_seed(cin + cout)
    n, h, w = 8, 120, 130
    x = ops.as_act(torch.randn(n, cin, h, w, generator=g).cuda().to(torch.bfloat16))
    res = ops.as_act(torch.randn(n, cout, h, w, generator=g).cuda().to(torch.bfloat16))
    wt = (torch.randn(cout, k, k, cin, generator=g) * (2.0 / (cin * k * k)) ** 0.5).cuda().to(torch.bfloat16)
    pw = ops.PackedConv(wt, 1)
    bias = torch.randn(cout, generator=g).cuda()
    ybuf = ops.as_act(torch.zeros(n, cout + 16, h, w).cuda().to(torch.bfloat16))
    y0 = ops.conv2d(x, pw, bias, k, 1, act="silu", impl=2, residual=res, out=ybuf[:, 8:8 + cout])
    y1 = ops.conv2d(x, pw, bias, k, 1, act="silu", impl=1, residual=res)
    torch.cuda.synchronize()
    scale = float(y1.float().abs().max())
    err = float((y0.float() - y1.float()).abs().max()) / scale
    assert err <= 2 ** -7, f"{err:.3e}"
    assert float(ybuf[:, :8].float().abs().max()) == 0.0 and float(ybuf[:, 8 + cout:].float().abs().max()) == 0.0


@pytest.mark.parametrize("shape", [(3, 32, 13, 17), (2, 32, 40, 40)])
def test_dcn_fused_groupnorm_stats(shape):
    """DyDCNv2: GroupNorm sums accumulated in the DCN conv's epilogue against the stand-alone statistics pass."""
    from mgdt_yolo_b200 import ops
    from mgdt_yolo_b200.modules.block import DyDCNv2
    from mgdt_yolo_b200.synth import synth_state_dict
    n, c, h, w = shape
    torch.manual_seed(h)
    m = DyDCNv2(c, c)
    m.load_state_dict(synth_state_dict(m.state_dict(), seed=w))
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(h * w)
    x = ops.as_act(torch.randn(n, c, h, w, generator=g).cuda().to(torch.bfloat16))
    off = ops.as_act((torch.randn(n, 18, h, w, generator=g) * 1.5).cuda().to(torch.bfloat16))
    msk = ops.as_act(torch.rand(n, 9, h, w, generator=g).cuda().to(torch.bfloat16))
    outs = {}
    for rnd in range(2):
        for fused in (True, False):
            ops.FUSE_STATS = fused
            try:
                with torch.no_grad():
                    y = m(x, off, msk, act="relu").float()
                torch.cuda.synchronize()
            finally:
                ops.FUSE_STATS = True
            if rnd and fused:
                assert torch.equal(y, outs[True])
            outs[fused] = y
    err = float((outs[True] - outs[False]).abs().max()) / float(outs[False].abs().max())
    assert err <= 2 ** -7, f"{err:.3e}"


@pytest.mark.parametrize("shape", [(384, 96, 1, 3, 13, 17), (384, 96, 1, 4, 40, 40), (64, 32, 1, 3, 80, 80), (64, 32, 3, 2, 21, 19),
                                   (32, 48, 1, 5, 8, 8)])
def test_per_image_weights_vs_in_scale(shape):
    """Per-(n, c) input scales folded into per-image weights (mgdt_conv_umma_pack_scaled + w_per_image, tiles cut per
    image, weight slices through the ring) against the in-loader activation transform and an fp32 reference; images
    whose pixel count is not a multiple of the tile, K-sliced layers, a residual and a channel-slice output."""
    from mgdt_yolo_b200 import ops
    cin, cout, k, n, h, w = shape
    g = torch.Generator().manual_seed(cin + h)
    x = ops.as_act(torch.randn(n, cin, h, w, generator=g).cuda().to(torch.bfloat16))
    res = ops.as_act(torch.randn(n, cout, h, w, generator=g).cuda().to(torch.bfloat16))
    wt32 = (torch.randn(cout, k, k, cin, generator=g) * (2.0 / (cin * k * k)) ** 0.5).cuda()
    pw = ops.PackedConv(wt32.to(torch.bfloat16), 1, w32=wt32)
    insc = (torch.rand(n, cin, generator=g) * 1.5 + 0.25).cuda().contiguous()
    bias = torch.randn(cout, generator=g).cuda()
    outs = {}
    for per_image in (True, False):
        ops.PER_IMAGE_WEIGHTS = per_image
        try:
            ybuf = ops.as_act(torch.zeros(n, cout + 16, h, w).cuda().to(torch.bfloat16))
            ops.conv2d(x, pw, bias, k, 1, act="silu", in_scale=insc, residual=res, out=ybuf[:, 8:8 + cout])
            torch.cuda.synchronize()
        finally:
            ops.PER_IMAGE_WEIGHTS = True
        assert float(ybuf[:, :8].float().abs().max()) == 0.0 and float(ybuf[:, 8 + cout:].float().abs().max()) == 0.0
        outs[per_image] = ybuf[:, 8:8 + cout].float()
    ref = F.silu(F.conv2d(x.float() * insc.view(n, cin, 1, 1), wt32.permute(0, 3, 1, 2), bias, padding=k // 2)) + res.float()
    scale = float(ref.abs().max())
    for key, y in outs.items():
        err = float((y - ref).abs().max()) / scale
        assert err <= 1e-2, f"per_image={key}: {err:.3e}"
    assert float((outs[True] - outs[False]).abs().max()) / scale <= 1e-2


@pytest.mark.parametrize("cout,n,h,w", [(2, 2, 17, 19), (2, 8, 120, 130), (48, 3, 20, 24)])
def test_pixel_scale_1x1_in_epilogue(cout, n, h, w):
    """cv3(cls_feat * cls_prob) (head.py:528): for a 1x1 conv the per-pixel input scale is applied to the accumulator row
    in the epilogue (W (s x) = s (W x)), also in the paired-unit form (Cout <= 16, large maps), against the CUDA-core
    path that scales the input."""
    from mgdt_yolo_b200 import ops
    cin = 32
    g = torch.Generator().manual_seed(cout + h)
    x = ops.as_act(torch.randn(n, cin, h, w, generator=g).cuda().to(torch.bfloat16))
    pix = ops.as_act(torch.rand(n, 1, h, w, generator=g).cuda().to(torch.bfloat16))
    wt = (torch.randn(cout, 1, 1, cin, generator=g) * (2.0 / cin) ** 0.5).cuda().to(torch.bfloat16)
    pw = ops.PackedConv(wt, 1)
    bias = torch.randn(cout, generator=g).cuda()
    y0 = ops.conv2d(x, pw, bias, 1, 1, pix_scale=pix, impl=2)
    y1 = ops.conv2d(x, pw, bias, 1, 1, pix_scale=pix, impl=1)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float() * pix.float(), wt.float().permute(0, 3, 1, 2), bias)
    scale = float(ref.abs().max())
    assert float((y0.float() - ref).abs().max()) / scale <= 2 ** -7
    assert float((y0.float() - y1.float()).abs().max()) / scale <= 2 ** -6


def test_tood_sibling_convs_as_one_gemm():
    """cls_decomp / reg_decomp reduction convs (each with its own layer attention folded into per-image weights) and
    cls_prob_conv1 (unscaled) as ONE per-image-weight GEMM (mgdt_conv_umma_pack_scaled_groups) against the three
    separate launches: same operands, same bf16 rounding of the scaled weights -> same maps."""
    from mgdt_yolo_b200 import ops
    from mgdt_yolo_b200.modules import TOODHead
    from mgdt_yolo_b200.synth import synth_state_dict
    head = TOODHead(3, 64, (64,))
    head.load_state_dict(synth_state_dict(head.state_dict(), seed=5))
    head.stride = torch.tensor([8.0])
    head = head.cuda().eval()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(3, 64, 40, 56, generator=g).cuda().to(torch.bfloat16)
    outs = {}
    for flag in (True, False):
        ops.FUSE_TOOD_SIBLINGS = flag
        try:
            with torch.no_grad():
                y, raw = head([x.clone()])
        finally:
            ops.FUSE_TOOD_SIBLINGS = False
        outs[flag] = (y.float(), raw[0].float())
    for a, b in zip(outs[True], outs[False]):
        err = float((a - b).abs().max()) / float(b.abs().max())
        assert err <= 2 ** -7, f"fused sibling GEMM vs separate convs: {err:.3e}"


@pytest.mark.parametrize("c,n,h,w", [(8, 2, 21, 19), (8, 3, 160, 160), (16, 2, 16, 24), (16, 4, 80, 80), (32, 2, 20, 20),
                                     (32, 3, 40, 40), (32, 2, 80, 80), (16, 1, 3, 5), (8, 1, 1, 1), (32, 1, 7, 33)])
@pytest.mark.parametrize("act", ["silu", "relu"])
def test_conv3x3_warp_kernel(c, n, h, w, act):
    """The warp-level MMA kernel for 3x3 stride-1 layers with Cin = Cout in {8, 16, 32} (csrc/conv3x3_warp.cu, the impl=0
    dispatch of the Bottleneck pairs): against torch's fp32 conv2d on the identical bf16 operands (tolerance = bf16 output
    rounding), against the tcgen05 path it replaces, with a residual and channel-slice input / output / residual (the
    neighbouring channels of the buffers must stay untouched); ragged widths (tiles of 16 pixels), 1-pixel maps."""
    from mgdt_yolo_b200 import ops
    from mgdt_yolo_b200._lib import lib
    lib().mgdt_set_option(b"conv3x3_warp", 2)
    g = torch.Generator().manual_seed(c + h)
    xbuf = ops.as_act(torch.randn(n, 3 * c, h, w, generator=g).cuda().to(torch.bfloat16))
    x = xbuf[:, c:2 * c]
    wt = (torch.randn(c, 3, 3, c, generator=g) * (2.0 / (c * 9)) ** 0.5).cuda().to(torch.bfloat16)
    bias = (torch.randn(c, generator=g) * 0.1).cuda()
    pw = ops.PackedConv(wt, 1)
    ops.PROFILE = []
    try:
        obuf = ops.new_act(n, 3 * c, h, w, torch.bfloat16, "cuda")
        obuf.zero_()
        ops.conv2d(x, pw, bias, 3, 1, act=act, out=obuf[:, 2 * c:], residual=xbuf[:, :c])
        plain = ops.conv2d(x, pw, bias, 3, 1, act=act)
        kern = [m["kernel"] for _, m, _, _ in ops.PROFILE]
    finally:
        ops.PROFILE = None
        lib().mgdt_set_option(b"conv3x3_warp", 2)
    assert kern == ["conv3x3_warp_kernel"] * 2, kern
    try:
        umma = ops.conv2d(x, pw, bias, 3, 1, act=act, impl=2)
    except RuntimeError:
        umma = None          # degenerate maps the tcgen05 path refuses
    torch.cuda.synchronize()
    fa = F.silu if act == "silu" else F.relu
    ref = fa(F.conv2d(x.float(), wt.float().permute(0, 3, 1, 2), bias, padding=1))
    scale = max(float(ref.abs().max()), 1e-3)
    assert float(obuf[:, :2 * c].abs().max()) == 0
    assert float((plain.float() - ref).abs().max()) / scale <= 2 ** -7
    assert umma is None or float((plain.float() - umma.float()).abs().max()) / scale <= 2 ** -7
    ref_r = ref + xbuf[:, :c].float()
    assert float((obuf[:, 2 * c:].float() - ref_r).abs().max()) / max(float(ref_r.abs().max()), 1e-3) <= 2 ** -7


@pytest.mark.parametrize("c1,c2,nc,n,h,w", [(16, 32, 2, 3, 80, 80), (8, 16, 1, 2, 13, 9), (32, 64, 5, 2, 20, 21), (16, 32, 8, 1, 1, 1)])
def test_tood_cls_fused_tail(c1, c2, nc, n, h, w):
    """cv3(cls_feat * sigmoid(cls_prob_conv2(prob))) in one launch (csrc/tood_cls.cu) against the two convolutions it
    replaces (3x3 -> 1 channel + sigmoid, 1x1 with a per-pixel input scale) and against torch fp32 on the same bf16
    operands; writes only its nc channels of the raw map."""
    from mgdt_yolo_b200 import ops
    g = torch.Generator().manual_seed(c1 + nc)
    prob = ops.as_act(torch.rand(n, c1, h, w, generator=g).cuda().to(torch.bfloat16))
    featb = ops.as_act(torch.randn(n, 2 * c2, h, w, generator=g).cuda().to(torch.bfloat16))
    feat = featb[:, c2:]
    w2 = (torch.randn(1, 3, 3, c1, generator=g) * 0.2).cuda().to(torch.bfloat16)
    w3 = (torch.randn(nc, 1, 1, c2, generator=g) * 0.2).cuda().to(torch.bfloat16)
    b2, b3 = torch.randn(1, generator=g).cuda(), torch.randn(nc, generator=g).cuda()
    raw = ops.new_act(n, 64 + nc, h, w, torch.bfloat16, "cuda")
    raw.zero_()
    assert ops.tood_cls(prob, w2, b2, feat, w3, b3, raw[:, 64:]) is not None
    pr = ops.conv2d(prob, ops.PackedConv(w2, 1), b2, 3, act="sigmoid")
    two = ops.conv2d(feat, ops.PackedConv(w3, 1), b3, 1, pix_scale=pr)
    torch.cuda.synchronize()
    pr32 = torch.sigmoid(F.conv2d(prob.float(), w2.float().permute(0, 3, 1, 2), b2, padding=1)).to(torch.bfloat16).float()
    ref = F.conv2d((feat.float() * pr32).to(torch.bfloat16).float(), w3.float().permute(0, 3, 1, 2), b3)
    scale = max(float(ref.abs().max()), 1e-3)
    assert float(raw[:, :64].abs().max()) == 0
    assert float((raw[:, 64:].float() - ref).abs().max()) / scale <= 2 ** -6
    assert float((raw[:, 64:].float() - two.float()).abs().max()) / scale <= 2 ** -6



def test_conv3x3_warp_kernel_persistent_option():
    """`conv3x3_warp_spc` > 1: persistent CTAs walking several strips with double-buffered staging (off by default, no gain
    measured) must give the same bytes as one CTA per strip."""
    from mgdt_yolo_b200 import ops
    from mgdt_yolo_b200._lib import lib
    g = torch.Generator().manual_seed(21)
    outs = {}
    for c, n, h, w in ((8, 32, 160, 160), (16, 32, 80, 80), (32, 16, 40, 44)):
        x = ops.as_act(torch.randn(n, c, h, w, generator=g).cuda().to(torch.bfloat16))
        wt = (torch.randn(c, 3, 3, c, generator=g) * (2.0 / (c * 9)) ** 0.5).cuda().to(torch.bfloat16)
        bias = (torch.randn(c, generator=g) * 0.1).cuda()
        pw = ops.PackedConv(wt, 1)
        for spc in (1, 3):
            lib().mgdt_set_option(b"conv3x3_warp_spc", spc)
            try:
                outs[spc] = ops.conv2d(x, pw, bias, 3, 1, act="silu", residual=x)
            finally:
                lib().mgdt_set_option(b"conv3x3_warp_spc", 1)
        torch.cuda.synchronize()
        assert torch.equal(outs[1], outs[3]), f"persistent strips differ for C = {c}"
