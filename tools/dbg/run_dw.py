import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mgdt_yolo_b200 import ops
from mgdt_yolo_b200.modules.block import ConvNeXtV2_Block
from mgdt_yolo_b200.synth import synth_state_dict
m = ConvNeXtV2_Block(96)
m.load_state_dict(synth_state_dict(m.state_dict(), seed=3))
m = m.cuda().eval()
x = ops.as_act(torch.randn(32, 96, 40, 40, device="cuda").to(torch.bfloat16))
with torch.no_grad():
    for _ in range(3):
        y = m(x)
torch.cuda.synchronize()
