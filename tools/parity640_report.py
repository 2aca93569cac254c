"""Per-layer parity table of the four BASELINE configs at 640 x 640 against the CPU oracle (run on the GPU box):
fp32 validation mode at B = 2 and bf16 at B = 32 -> gpurun_out/parity640.json (profiles/parity640_rNN.md is derived)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests import parity  # noqa: E402

out = {}
cfgs = sys.argv[1:] or list(parity.BASELINE_CFGS)
for cfg in cfgs:
    for dtype, B in ((torch.float32, 2), (torch.bfloat16, 32)):
        res, _, _ = parity.compare_640(cfg, dtype, B)
        key = f"{cfg}|{str(dtype)[6:]}|B{B}"
        out[key] = {k: [float(f"{v[0]:.4e}"), float(f"{v[1]:.4e}")] for k, v in res.items()}
        worst = max(res.items(), key=lambda kv: kv[1][0])
        print(key, "worst max-rel", worst[0], f"{worst[1][0]:.3e}", " y", res["y"], flush=True)
        torch.cuda.empty_cache()
for cfg in cfgs:   # the yardstick: the reference's own modules in bf16 (torch eager) against the same oracle
    r = parity.reference_bf16_640(cfg, 32)
    if r is not None:
        out[f"{cfg}|reference-bf16-eager|B32"] = {k: [float(f"{v[0]:.4e}"), float(f"{v[1]:.4e}")] for k, v in r.items()}
        print(cfg, "reference in bf16: y", r["y"], flush=True)
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "parity640.json"), "w") as f:
    json.dump(out, f, indent=1)
