"""Isolated timing of the 3x3 stride-1 Bottleneck convs: warp-level MMA kernel (impl 0) vs the tcgen05 path (impl 2).
usage: python tools/dbg/cw_bench.py [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from mgdt_yolo_b200 import ops

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for c, h in ((8, 160), (16, 80), (32, 80), (32, 40)):
    n = 32
    xbuf = ops.as_act(torch.randn(n, 4 * c, h, h, generator=g).cuda().to(torch.bfloat16))
    x, out = xbuf[:, c:2 * c], xbuf[:, 2 * c:3 * c]
    wt = (torch.randn(c, 3, 3, c, generator=g) * (2.0 / (c * 9)) ** 0.5).cuda().to(torch.bfloat16)
    bias = torch.zeros(c, device="cuda")
    pw = ops.PackedConv(wt, 1)
    for impl in (0, 2):
        ts = []
        for i in range(iters + 3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv2d(x, pw, bias, 3, 1, act="silu", out=out, residual=x, impl=impl)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        print(f"{c}->{c} @{h}^2 impl {impl}: median {ts[len(ts) // 2]:.1f} us, min {ts[0]:.1f} us")
