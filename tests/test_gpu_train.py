"""Row f3 on the B200 (pytest -m gpu): the fused CUDA criterion (mgdt_v8_loss) against the live-reference fixture
tests/golden/loss.npz and against the reference run on the same GPU (baseline/_ref), the flat-bucket optimizer launches
against torch.optim.SGD + clip_grad_norm_ + the EMA formula, and a short training run of the full config.
Tolerance: 1e-4 relative (BASELINE.json north_star, fp32 mode)."""
import os

import numpy as np
import pytest
import torch

from baseline import ref_loader
from oracle.train_cases import LOSS_CASES, loss_inputs, model_stub

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss.npz")


def _ours(name, device="cuda"):
    from mgdt_yolo_b200.train import v8DetectionLoss
    b, nc, reg_max, levels, counts, seed, calls = LOSS_CASES[name]
    feats, batch = loss_inputs(name)
    feats = [f.to(device).requires_grad_(True) for f in feats]
    crit = v8DetectionLoss(model_stub(nc, reg_max, levels, device))
    crit.epoch = calls
    crit.keep_assignment = True
    total, items = crit(list(feats), {k: v.to(device) for k, v in batch.items()})
    total.backward()
    return total, items, feats, crit


def _rel(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-12)


@pytest.mark.parametrize("name", list(LOSS_CASES))
def test_criterion_vs_live_reference_fixture(name):
    g = np.load(GOLDEN)
    total, items, feats, crit = _ours(name)
    assert _rel(items, g[f"{name}.items"]) <= 1e-4, (items.tolist(), g[f"{name}.items"].tolist())
    assert _rel(total, g[f"{name}.total"]) <= 1e-4
    for i, f in enumerate(feats):
        assert _rel(f.grad, g[f"{name}.grad{i}"]) <= 1e-4, f"gradient of level {i}"
    # assigner: normalised target scores (B, A, nc) and the positives that carry weight
    ts_ref = torch.from_numpy(g[f"{name}.target_scores"])
    lab, ts = crit.last["target_labels"].cpu().long(), crit.last["target_scores"].cpu()
    dense = torch.zeros_like(ts_ref)
    pos = lab >= 0
    dense[pos, lab[pos]] = ts[pos]
    assert _rel(dense, ts_ref) <= 1e-4
    fg_ref = torch.from_numpy(g[f"{name}.fg_mask"]).bool()
    weighty = ts_ref.sum(-1) > 0
    assert torch.equal(pos & weighty, fg_ref & weighty)      # zero-metric positives depend on topk's tie order and carry no loss
    tb_ref = torch.from_numpy(g[f"{name}.target_bboxes"])      # the reference divides them by the stride in place (loss.py:196)
    # (anchors whose weight is ~0 may pick another box when two alignment metrics differ in the last bit: no effect on the loss)
    sel = pos & fg_ref & (ts_ref.sum(-1) > 1e-4 * float(ts_ref.max()))
    ours_tb = crit.last["target_bboxes"].detach().cpu() / crit._strides.cpu().view(1, -1, 1)
    assert _rel(ours_tb[sel], tb_ref[sel]) <= 1e-5


@pytest.mark.skipif(not ref_loader.available(), reason="reference copy (baseline/_ref) not present")
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_criterion_vs_reference_on_gpu(seed):
    """The reference's own v8DetectionLoss on this GPU, larger problem: B = 8, one 80 x 80 level (the TOOD head at
    640 x 640), ~20 labels per image."""
    from mgdt_yolo_b200.train import synth_targets, v8DetectionLoss
    ref_loader.load()
    from ultralytics.yolo.utils import loss as L
    b, nc, reg_max, levels = 8, 2, 16, [(80, 80, 8.0)]
    g = torch.Generator().manual_seed(100 + seed)
    f = torch.randn(b, nc + 4 * reg_max, 80, 80, generator=g)
    f[:, :64] *= 1.5
    batch = synth_targets(b, 20, nc, seed=seed, device="cuda")
    fa, fb = f.cuda().requires_grad_(True), f.cuda().requires_grad_(True)
    want, witems = L.v8DetectionLoss(model_stub(nc, reg_max, levels, "cuda"))([fa], batch)
    got, gitems = v8DetectionLoss(model_stub(nc, reg_max, levels, "cuda"))([fb], batch)
    want.backward()
    got.backward()
    assert _rel(gitems, witems) <= 1e-4 and _rel(got, want) <= 1e-4
    assert _rel(fb.grad, fa.grad) <= 1e-4


def test_criterion_without_targets():
    from mgdt_yolo_b200.train import v8DetectionLoss
    nc, reg_max, levels = 2, 16, [(8, 8, 8.0)]
    f = torch.randn(2, 66, 8, 8, device="cuda", requires_grad=True)
    batch = {"cls": torch.zeros(0, 1), "bboxes": torch.zeros(0, 4), "batch_idx": torch.zeros(0)}
    total, items = v8DetectionLoss(model_stub(nc, reg_max, levels, "cuda"))([f], batch)
    total.backward()
    want = torch.nn.functional.binary_cross_entropy_with_logits(f[:, 64:], torch.zeros_like(f[:, 64:]), reduction="sum") * 0.5
    assert items[0] == 0 and items[2] == 0 and _rel(items[1], want.detach()) <= 1e-5
    assert float(f.grad[:, :64].abs().max()) == 0.0


def test_flat_bucket_step_matches_torch():
    """mgdt_sumsq + mgdt_sgd_step + mgdt_ema_update against clip_grad_norm_(10) + SGD(nesterov, three groups) + the
    ModelEMA recurrence, three steps."""
    import copy
    import math
    from mgdt_yolo_b200.modules import C2f
    from mgdt_yolo_b200.train import FlatBucket
    torch.manual_seed(0)
    m = C2f(8, 8, 1).cuda().train()
    ref = copy.deepcopy(m)
    bias, norm, rest = [], [], []
    for mod in ref.modules():
        for k, p in mod.named_parameters(recurse=False):
            (bias if "bias" in k else norm if "Norm" in type(mod).__name__ else rest).append(p)
    opt = torch.optim.SGD([{"params": bias, "weight_decay": 0.0}, {"params": norm, "weight_decay": 0.0},
                           {"params": rest, "weight_decay": 5e-4}], lr=0.01, momentum=0.937, nesterov=True)
    ema = {k: v.detach().clone() for k, v in ref.named_parameters()}
    fb = FlatBucket(m, lr=0.01, momentum=0.937, weight_decay=5e-4)
    for it in range(3):
        x = torch.randn(4, 8, 16, 16, device="cuda") * (30.0 if it == 0 else 1.0)     # first step: the clip is active
        fb.zero_grad()
        (m(x) ** 2).sum().backward()
        fb.step(fb.all_reduce())
        opt.zero_grad()
        (ref(x) ** 2).sum().backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm=10.0)
        opt.step()
        d = 0.9999 * (1 - math.exp(-(it + 1) / 2000))
        for k, p in ref.named_parameters():
            ema[k].mul_(d).add_(p.detach(), alpha=1 - d)
    for (k, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        assert _rel(p.detach(), q.detach()) <= 1e-5, k
    sd = fb.ema_state_dict()
    for k in ema:
        assert _rel(sd[k], ema[k]) <= 1e-5, k


def test_train_steps_full_config_then_inference():
    """Three training steps of the full MGDT config (bf16 autocast forward / backward, CUDA criterion, fused optimizer
    side), then the SAME module objects in eval mode on the inference kernels (weights repacked after the update)."""
    from mgdt_yolo_b200.synth import synth_images, synth_state_dict
    from mgdt_yolo_b200.tasks import DetectionModel
    from mgdt_yolo_b200.train import FlatBucket, synth_targets, train_step
    m = DetectionModel("mspa_c2f_gd_tood_yolov8n.yaml", nc=2, verbose=False)
    m.load_state_dict(synth_state_dict(m.state_dict(), seed=1))
    m = m.cuda()
    fb = FlatBucket(m, lr=0.01)
    batch = synth_targets(4, 10, 2, seed=0)
    batch["img"] = synth_images(4, h=128, w=160, seed=0).cuda()
    before = fb.flat.clone()
    losses = []
    for _ in range(3):
        loss, items = train_step(m, fb, batch)
        assert torch.isfinite(loss) and torch.isfinite(items).all()
        losses.append(float(loss))
    assert float((fb.flat - before).abs().max()) > 0
    assert float((fb.ema - before).abs().max()) > 0
    m.eval()
    with torch.no_grad():
        y, raw = m(batch["img"].to(torch.bfloat16))
    assert y.shape == (4, 6, 16 * 20) and torch.isfinite(y.float()).all()


@pytest.mark.skipif(not ref_loader.available(), reason="reference copy (baseline/_ref) not present")
def test_model_loss_and_gradients_match_reference_model():
    """`model(batch)` as the trainer calls it (trainer.py:329-334): the drop-in DetectionModel (train-mode modules + the fused
    CUDA criterion) against the reference's DetectionModel + its own v8DetectionLoss on the same GPU, fp32, TF32 off."""
    import types
    from mgdt_yolo_b200.synth import synth_images, synth_state_dict
    from mgdt_yolo_b200.tasks import DetectionModel
    from mgdt_yolo_b200.train import synth_targets
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = "mspa_c2f_gd_tood_yolov8n.yaml"
    ref = ref_loader.build_model(cfg, nc=2)
    sd = synth_state_dict(ref.state_dict(), seed=1)
    ref.load_state_dict(sd)
    ours = DetectionModel(cfg, nc=2, verbose=False)
    ours.load_state_dict(sd)
    hyp = types.SimpleNamespace(box=7.5, cls=0.5, dfl=1.5)
    ref, ours = ref.cuda().train(), ours.cuda().train()
    ref.args, ours.args = hyp, hyp
    batch = synth_targets(4, 10, 2, seed=3)
    batch["img"] = synth_images(4, h=128, w=160, seed=2).cuda()
    lr, ir = ref(dict(batch))
    lo, io = ours(dict(batch))
    assert _rel(lo, lr) <= 1e-4 and _rel(io, ir) <= 1e-4, (io.tolist(), ir.tolist())
    lr.backward()
    lo.backward()
    pr, po = dict(ref.named_parameters()), dict(ours.named_parameters())
    gmax = max(float(p.grad.abs().max()) for p in pr.values() if p.grad is not None)
    for k, p in pr.items():
        if p.grad is None:
            assert po[k].grad is None or float(po[k].grad.abs().max()) == 0.0, k
        else:   # atomics in cuDNN / deform_conv2d backward: compare on the scale of the largest gradient
            assert float((po[k].grad - p.grad).abs().max()) <= 2e-3 * max(float(p.grad.abs().max()), 1e-3 * gmax), k
