"""profiles/parity640_rNN.md: per-layer deviation from the fp32 oracle at 640 x 640 of (a) this repo's bf16 path and (b) the
REFERENCE's own modules run in bf16 by torch eager on the same GPU (baseline/_ref) -- the yardstick behind the bf16
tolerance of tests/parity.py.  Run on the GPU box:  python tools/parity_table.py gpurun_out/parity640_r02.md [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests import parity  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/parity640.md"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg = "mspa_c2f_gd_tood_yolov8n.yaml"
ours, _, _ = parity.compare_640(cfg, torch.bfloat16, B)
ours32, _, _ = parity.compare_640(cfg, torch.float32, 2)
ref = parity.reference_bf16_640(cfg, B) or {}
with open(out, "w") as f:
    f.write(f"# Deviation from the fp32 oracle at 640 x 640, full MGDT config ({cfg}), B = {B} (fp32 column: B = 2)\n\n"
            "max-rel = max|a - b| / max|b|, L2 = ||a - b|| / ||b|| against `oracle.forward` on identical weights and inputs.\n"
            "`reference bf16` = the unmodified reference modules (`baseline/_ref`, fused BN) in bf16 through torch eager / cuDNN on the\n"
            "same B200: the format's own error on this graph.\n\n"
            "| tensor | ours fp32 max-rel | ours bf16 max-rel | ours bf16 L2 | reference bf16 max-rel | reference bf16 L2 |\n|---|---:|---:|---:|---:|---:|\n")
    for k, (mx, l2) in ours.items():
        r = ref.get(k)
        f32 = ours32.get(k, (float("nan"),))[0]
        f.write(f"| {k} | {f32:.1e} | {mx:.2e} | {l2:.2e} | " + (f"{r[0]:.2e} | {r[1]:.2e} |\n" if r else "- | - |\n"))
print(open(out).read())
