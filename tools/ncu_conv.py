"""Small driver for ncu: a handful of representative launches of the persistent tcgen05 conv kernel
(B=32 shapes of the full MGDT config), each after one warm-up launch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mgdt_yolo_b200 import ops  # noqa: E402
from mgdt_yolo_b200.modules import Conv  # noqa: E402
from mgdt_yolo_b200.synth import synth_state_dict  # noqa: E402

torch.manual_seed(0)
B = 32
cases = [  # cin, cout, k, s, H, W, act
    ("stem", 3, 16, 3, 2, 640, 640),
    ("c8_1x1", 8, 8, 1, 1, 160, 160),
    ("c32_3x3", 32, 32, 3, 1, 80, 80),
    ("c16_32_s2", 16, 32, 3, 2, 320, 320),
    ("c96_384", 96, 384, 1, 1, 40, 40),
    ("c64_256", 64, 256, 1, 1, 80, 80),
    ("c256_64", 256, 64, 1, 1, 80, 80),
    ("c32_64", 32, 64, 1, 1, 80, 80),
    ("c64_256_noact", 64, 256, 1, 1, 80, 80),
]
sel = sys.argv[1:] or [c[0] for c in cases]
for name, cin, cout, k, s, H, W in cases:
    if name not in sel:
        continue
    m = Conv(cin, cout, k, s, act=not name.endswith('_noact'))
    m.load_state_dict(synth_state_dict(m.state_dict(), seed=3))
    m = m.cuda().eval()
    with torch.no_grad():
        if name == "stem":
            src = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, device="cuda")
            for _ in range(2):
                y = m.forward_image(src)
        else:
            x = ops.as_act(torch.randn(B, cin, H, W, device="cuda").to(torch.bfloat16))
            for _ in range(2):
                y = m(x)
    torch.cuda.synchronize()
    print(name, tuple(y.shape), float(y.float().abs().mean()))
