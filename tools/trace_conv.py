"""Phase timeline of the persistent tcgen05 conv kernel for a few shapes (debug tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mgdt_yolo_b200 import ops
from mgdt_yolo_b200._lib import lib
from mgdt_yolo_b200.modules import Conv
from mgdt_yolo_b200.synth import synth_state_dict

B = 32
import torch.nn as nn
cases = [("c32_3x3", 32, 32, 3, 1, 80, 80), ("c8_1x1", 8, 8, 1, 1, 160, 160), ("c64_256", 64, 256, 1, 1, 80, 80),
         ("c64_256_noact", 64, 256, 1, 1, 80, 80), ("c64_256_relu", 64, 256, 1, 1, 80, 80),
         ("c96_384", 96, 384, 1, 1, 40, 40), ("c16_3x3", 16, 16, 3, 1, 80, 80), ("c32_64", 32, 64, 1, 1, 80, 80),
         ("c256_64", 256, 64, 1, 1, 80, 80), ("c16_32_s2", 16, 32, 3, 2, 320, 320), ("c384_96", 384, 96, 1, 1, 40, 40), ("c384_96_grn", 384, 96, 1, 1, 40, 40)]
cases += [("c96_384_gelu", 96, 384, 1, 1, 40, 40), ("c96_384_gelu_sq", 96, 384, 1, 1, 40, 40), ("c96_384_silu_sq", 96, 384, 1, 1, 40, 40),
          ("c32_32_q5", 32, 32, 1, 1, 160, 160), ("c32_32_160", 32, 32, 1, 1, 160, 160)]
cases += [("c32_3x3_w78", 32, 32, 3, 1, 80, 78), ("c32_3x3_w86", 32, 32, 3, 1, 80, 86), ("c32_3x3_w62", 32, 32, 3, 1, 80, 62)]
cases.append(("stem", 3, 16, 3, 2, 640, 640))
cases += [("c8_3x3", 8, 8, 3, 1, 160, 160), ("c64_32_3x3", 64, 32, 3, 1, 80, 80), ("c32_3x3_40", 32, 32, 3, 1, 40, 40)]
sel = sys.argv[1:]
if sel:
    cases = [c for c in cases if c[0] in sel]
trace = torch.zeros(148 * 16 * 64, dtype=torch.int64, device="cuda")
for name, cin, cout, k, s, H, W in cases:
    m = Conv(cin, cout, k, s, act=(False if 'noact' in name else nn.ReLU() if 'relu' in name else nn.GELU() if 'gelu' in name else True))
    m.load_state_dict(synth_state_dict(m.state_dict(), seed=3))
    m = m.cuda().eval()
    x = ops.as_act(torch.randn(B, cin, H, W, device="cuda").to(torch.bfloat16))
    if name == "stem":   # fused uint8 preprocessing + layer-0 conv
        x = torch.randint(0, 256, (B, cin, H, W), dtype=torch.uint8, device="cuda")
        m.forward = m.forward_image
    with torch.no_grad():
        kw = {}
        if name.endswith("_grn"):   # GRN-scaled input + residual, as ConvNeXtV2_Block's pwconv2
            kw = dict(in_scale=(torch.rand(B, cin, device="cuda") + 0.5).contiguous(),
                      residual=ops.as_act(torch.randn(B, cout, H, W, device="cuda").to(torch.bfloat16)))
        if name.endswith("_sq"):
            kw = dict(stat=ops.StatReq(0, True))
        if name.endswith("_q5"):
            kw = dict(stat=ops.StatReq(5, False))
        for _ in range(3):
            m(x, **kw)
        torch.cuda.synchronize()
        trace.zero_()
        lib().mgdt_debug_set_trace(trace.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); m(x, **kw); e1.record()
        torch.cuda.synchronize()
        lib().mgdt_debug_set_trace(None)
    t = trace.view(-1, 64).cpu()
    act = t[:, 0] > 0
    t = t[act]
    if len(t) == 0:
        print(f"== {name}: event time {e0.elapsed_time(e1)*1000:.1f} us -- not the tcgen05 kernel (no trace)")
        continue
    t0 = int(t[:, 0].min())
    rel = lambda col: [(int(v) - t0) / 1000 for v in col if v > 0]
    def stat(slot):
        r = rel(t[:, slot])
        return f"{min(r):7.1f}/{sum(r)/len(r):7.1f}/{max(r):7.1f}" if r else "   -"
    print(f"== {name}: event time {e0.elapsed_time(e1)*1000:.1f} us, CTAs {len(t)}   (us since first CTA start: min/mean/max)")
    print("   start", stat(0), " setup", stat(1), " W", stat(2), " end", stat(3))
    for ti in range(6):
        if (t[:, 8 + 8 * ti] > 0).any():
            print(f"   tile{ti}: fill {stat(8+8*ti)}  mma_issued {stat(9+8*ti)}  acc_ready {stat(10+8*ti)}  epi_done {stat(11+8*ti)}")
            if (t[:, 12 + 8 * ti] > 0).any():
                print(f"          xform: copies of the previous item landed {stat(12+8*ti)}  transformed {stat(14+8*ti)}  published {stat(13+8*ti)}")
    ph = t[:, 56:60].float()
    n = ph[:, 0].clamp(min=1)
    print(f"   epilogue unit phases (cycles, warp 8, tile 1): units {ph[:,0].mean():.1f}  tmem-ld {(ph[:,1]/n).mean():.0f}"
          f"  math+stage {(ph[:,2]/n).mean():.0f}  store {(ph[:,3]/n).mean():.0f}")
