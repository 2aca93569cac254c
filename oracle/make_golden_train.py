"""TEST INFRASTRUCTURE -- golden vectors of the training criterion (row f3), from the LIVE reference.

Runs the unmodified v8DetectionLoss (yolo/utils/loss.py:108-208) with its assigners (yolo/utils/tal.py) on the synthetic
cases of oracle/train_cases.py and stores the loss items, the gradient with respect to the head output and the
assigner's results in tests/golden/loss.npz.  The CUDA criterion (mgdt_v8_loss) is held to these on the GPU box.

    python oracle/make_golden_train.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_live  # noqa: E402
from oracle.train_cases import LOSS_CASES, loss_inputs, model_stub  # noqa: E402


def run_reference(name, device="cpu"):
    """-> dict of numpy arrays: total, items, grad<i>, target_scores, fg_mask, target_bboxes, target_gt_idx"""
    ref_live.load()
    from ultralytics.yolo.utils import loss as L
    b, nc, reg_max, levels, counts, seed, calls = LOSS_CASES[name]
    feats, batch = loss_inputs(name)
    feats = [f.to(device).requires_grad_(True) for f in feats]
    batch = {k: v.to(device) for k, v in batch.items()}
    crit = L.v8DetectionLoss(model_stub(nc, reg_max, levels, device))
    crit.epoch = calls
    captured = {}
    assign = crit.assigner.task_aligned_assigner

    def hook(mod, args, out):
        captured.update(target_labels=out[0], target_bboxes=out[1], target_scores=out[2], fg_mask=out[3], target_gt_idx=out[4])
    h = assign.register_forward_hook(hook)
    total, items = crit(list(feats), batch)
    h.remove()
    total.backward()
    blob = {"total": total.detach().cpu().numpy(), "items": items.cpu().numpy()}
    for i, f in enumerate(feats):
        blob[f"grad{i}"] = f.grad.cpu().numpy()
    blob["target_scores"] = captured["target_scores"].float().cpu().numpy()
    blob["fg_mask"] = captured["fg_mask"].cpu().numpy()
    blob["target_bboxes"] = captured["target_bboxes"].float().cpu().numpy()
    blob["target_gt_idx"] = captured["target_gt_idx"].cpu().numpy()
    return blob


def main():
    out = {}
    for name in LOSS_CASES:
        for k, v in run_reference(name).items():
            out[f"{name}.{k}"] = v
        print(name, "items", out[f"{name}.items"], "positives", int(out[f"{name}.fg_mask"].sum()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "loss.npz"), **out)


if __name__ == "__main__":
    main()
