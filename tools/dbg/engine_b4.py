import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tests import parity
from mgdt_yolo_b200.engine import Engine
from mgdt_yolo_b200 import ops
FULL = "mspa_c2f_gd_tood_yolov8n.yaml"
m, sd = parity.build_model(FULL, nc=2, cls_bias=-1.238)
for B in (1, 2, 4, 8):
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (B, 3, 640, 640), dtype=torch.uint8, generator=g)
    x32 = (u8.float() / 255).cuda()
    with torch.no_grad():
        y32, _ = m(x32)
        y16, _ = m(x32.to(torch.bfloat16))
    eng = Engine(m, B, 640, torch.bfloat16, "cuda:0", conf=0.25, iou=0.7, slots=1)
    dets = eng(u8.pin_memory())
    pred = eng.slots[0].pred.float()
    print(B, "eager bf16 vs fp32", parity.errs(y16, y32), "engine vs fp32", parity.errs(pred, y32), "ndet", [d.shape[0] for d in dets])
    dets2 = eng(u8.pin_memory())
    print("   second call", parity.errs(eng.slots[0].pred.float(), y32))
    if B == 4:
        rng = np.random.default_rng(11)
        ims = [rng.integers(0, 256, (480, 640, 3), dtype=np.uint8) for _ in range(4)]
        r = eng.predict(ims, auto=False)
        print("   predict ndet", [x.boxes.data.shape[0] for x in r], "pred vs ?", float(eng.slots[0].pred.float().abs().max()))
        src = eng.slots[0].src.clone()
        with torch.no_grad():
            y32b, _ = m(src.float() / 255)
        print("   predict-path engine vs fp32 on same src", parity.errs(eng.slots[0].pred.float(), y32b))
