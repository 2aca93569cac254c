"""Phase timeline of the DCNv2 launch of the TOOD head (conv_umma2_kernel, LOADER 2: the producers build the sampled
im2col); same trace slots as tools/trace_conv.py."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mgdt_yolo_b200 import ops
from mgdt_yolo_b200._lib import lib
from mgdt_yolo_b200.modules.block import DyDCNv2

B, C, H, W = 32, 32, 80, 80
dm = DyDCNv2(C, C).cuda().eval()
x = ops.as_act(torch.randn(B, C, H, W, device="cuda").to(torch.bfloat16))
off = ops.as_act((torch.randn(B, 18, H, W, device="cuda") * 2).to(torch.bfloat16))
msk = ops.as_act(torch.randn(B, 9, H, W, device="cuda").to(torch.bfloat16))
trace = torch.zeros(148 * 16 * 64, dtype=torch.int64, device="cuda")
with torch.no_grad():
    for _ in range(3):
        dm(x, off, msk, mask_is_logit=True, act="relu")
    torch.cuda.synchronize()
    lib().mgdt_debug_set_trace(trace.data_ptr())
    dm(x, off, msk, mask_is_logit=True, act="relu")
    torch.cuda.synchronize()
    lib().mgdt_debug_set_trace(None)
t = trace.view(-1, 64).cpu()
t = t[t[:, 0] > 0]
t0 = int(t[:, 0].min())
rel = lambda col: [(int(v) - t0) / 1000 for v in col if v > 0]
def stat(slot):
    r = rel(t[:, slot])
    return f"{min(r):7.1f}/{sum(r)/len(r):7.1f}/{max(r):7.1f}" if r else "   -"
print(f"== dcn {C}->{C} {B}x{H}x{W}: CTAs {len(t)}   (us since first CTA start: min/mean/max)")
print("   start", stat(0), " setup", stat(1), " W", stat(2), " end", stat(3))
for ti in range(6):
    if (t[:, 8 + 8 * ti] > 0).any():
        print(f"   tile{ti}: fill {stat(8+8*ti)}  mma_issued {stat(9+8*ti)}  acc_ready {stat(10+8*ti)}  epi_done {stat(11+8*ti)}")
