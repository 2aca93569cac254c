"""TEST INFRASTRUCTURE -- synthetic inputs of the training-criterion parity cases (row f3), shared by the fixture
generator (oracle/make_golden_train.py, live reference) and the tests (CUDA criterion).  CPU generators only, so both
sides see identical numbers."""
from __future__ import annotations

import types

import torch

# name: (batch, nc, reg_max, [(h, w, stride) per level], labels per image, seed, criterion calls before this one)
LOSS_CASES = {
    "tood_1level": (3, 2, 16, [(20, 24, 8.0)], [7, 0, 12], 11, 0),
    "detect_3level": (2, 5, 4, [(16, 16, 8.0), (8, 8, 16.0), (4, 4, 32.0)], [9, 4], 12, 0),
    "tood_annealed": (2, 2, 16, [(12, 20, 8.0)], [15, 15], 13, 161 * 7),      # alpha = 0.5 * 93 / 100 (tal.py:266)
    "crowded": (2, 3, 16, [(10, 10, 8.0)], [40, 25], 14, 0),                  # anchors claimed by several boxes
}


def head_stub(nc, reg_max, levels):
    """What v8DetectionLoss.__init__ reads from model.model[-1] (loss.py:116-124)."""
    return types.SimpleNamespace(nc=nc, reg_max=reg_max, no=nc + 4 * reg_max, stride=torch.tensor([s for _, _, s in levels]))


def model_stub(nc, reg_max, levels, device="cpu"):
    p = torch.zeros(1, device=device)
    return types.SimpleNamespace(args=types.SimpleNamespace(box=7.5, cls=0.5, dfl=1.5), model=[head_stub(nc, reg_max, levels)],
                                 parameters=lambda: iter([p]))


def loss_inputs(name):
    """-> (feats list of (B, no, h, w) fp32, batch dict) on the CPU."""
    b, nc, reg_max, levels, counts, seed, _ = LOSS_CASES[name]
    g = torch.Generator().manual_seed(seed)
    no = nc + 4 * reg_max
    feats = []
    for h, w, _ in levels:
        f = torch.randn(b, no, h, w, generator=g)
        f[:, :4 * reg_max] *= 1.5                       # peaked DFL distributions: boxes of varied sizes
        f[:, 4 * reg_max:] = f[:, 4 * reg_max:] * 2 - 1
        feats.append(f)
    n = sum(counts)
    cxy = torch.rand(n, 2, generator=g) * 0.8 + 0.1
    wh = torch.rand(n, 2, generator=g) * 0.45 + 0.08
    batch = {"cls": torch.randint(0, nc, (n, 1), generator=g).float(), "bboxes": torch.cat([cxy, wh], 1),
             "batch_idx": torch.cat([torch.full((c,), float(i)) for i, c in enumerate(counts)])}
    return feats, batch
