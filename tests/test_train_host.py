"""Row f3 on the CPU: the train-mode forward / backward of the drop-in modules against the live reference, the
criterion's host side, the flat bucket and its all-reduce on two gloo ranks.  (The CUDA criterion and the fused
optimizer launches are checked on the B200 box: tests/test_gpu_train.py.)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from baseline import ref_loader
from mgdt_yolo_b200.synth import synth_images, synth_state_dict
from mgdt_yolo_b200.tasks import DetectionModel

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout / baseline/_ref not present")


@needs_ref
@pytest.mark.parametrize("cfg", ["mspa_c2f_gd_tood_yolov8n.yaml", "yolov8n.yaml", "mspa_c2f_gd_yolov8n.yaml"])
def test_train_mode_forward_backward_matches_reference(cfg):
    """Same weights, same input, both in train mode (batch-statistics BatchNorm): raw head maps, every parameter
    gradient and the updated running statistics agree (fp32, 1e-5 of the tensor's magnitude)."""
    ref = ref_loader.build_model(cfg)
    sd = synth_state_dict(ref.state_dict(), seed=1)
    ref.load_state_dict(sd)
    ref.train()
    ours = DetectionModel(cfg, verbose=False)
    ours.load_state_dict(sd)
    ours.train()
    x = synth_images(2, h=64, w=96, seed=0)
    a, b = ref(x.clone()), ours(x.clone())
    assert isinstance(b, list) and len(a) == len(b)          # train mode: the per-level raw maps (head.py:161-164, 530-533)
    for u, v in zip(a, b):
        assert u.shape == v.shape
        assert float((u - v).abs().max()) <= 1e-5 * float(u.abs().max())
    sum((t ** 2).sum() for t in a).backward()
    sum((t ** 2).sum() for t in b).backward()
    pa, pb = dict(ref.named_parameters()), dict(ours.named_parameters())
    assert pa.keys() == pb.keys()
    for k in pa:
        if pa[k].grad is None:                               # TOODHead.scale, reduction_conv biases
            assert pb[k].grad is None or float(pb[k].grad.abs().max()) == 0.0, k
            continue
        assert float((pa[k].grad - pb[k].grad).abs().max()) <= 1e-5 * max(float(pa[k].grad.abs().max()), 1e-12), k
    ba, bb = dict(ref.named_buffers()), dict(ours.named_buffers())
    for k in ba:
        assert torch.allclose(ba[k].float(), bb[k].float(), rtol=1e-5, atol=1e-7), k


@needs_ref
def test_preprocess_targets_matches_reference():
    from mgdt_yolo_b200.train import v8DetectionLoss
    from oracle.train_cases import LOSS_CASES, loss_inputs, model_stub
    ref_loader.load()
    from ultralytics.yolo.utils import loss as L
    for name, (b, nc, reg_max, levels, counts, seed, calls) in LOSS_CASES.items():
        _, batch = loss_inputs(name)
        stub = model_stub(nc, reg_max, levels)
        targets = torch.cat((batch["batch_idx"].view(-1, 1), batch["cls"].view(-1, 1), batch["bboxes"]), 1)
        scale = torch.tensor([640.0, 480.0, 640.0, 480.0])
        want = L.v8DetectionLoss(stub).preprocess(targets.clone(), b, scale)
        got = v8DetectionLoss(stub).preprocess(targets.clone(), b, scale)
        assert want.shape == got.shape and torch.allclose(want, got, rtol=0, atol=1e-4), name


def test_make_anchors():
    from mgdt_yolo_b200.train import make_anchors
    feats = [torch.zeros(1, 4, 3, 5), torch.zeros(1, 4, 2, 2)]
    pts, st = make_anchors(feats, [8.0, 16.0])
    assert pts.shape == (19, 2) and st.shape == (19, 1)
    assert pts[0].tolist() == [0.5, 0.5] and pts[4].tolist() == [4.5, 0.5] and pts[5].tolist() == [0.5, 1.5]
    assert st[14, 0] == 8.0 and st[15, 0] == 16.0


def test_model_has_training_surface():
    """forward(dict) -> loss (tasks.py:42-44, 204-216) and init_criterion exist; heads return lists in train mode."""
    m = DetectionModel("mspa_c2f_gd_tood_yolov8n.yaml", nc=2, verbose=False).train()
    out = m(torch.rand(1, 3, 64, 64))
    assert isinstance(out, list) and out[0].shape == (1, 66, 8, 8)
    assert callable(m.loss) and callable(m.init_criterion)


def test_flat_bucket_views_and_groups():
    from mgdt_yolo_b200.train import FlatBucket
    m = DetectionModel("mspa_c2f_gd_tood_yolov8n.yaml", nc=2, verbose=False).train()
    ref = {k: v.detach().clone() for k, v in m.named_parameters()}
    fb = FlatBucket(m)
    assert fb.n == sum(p.numel() for p in m.parameters() if p.requires_grad)
    for k, p in m.named_parameters():
        if p.requires_grad:
            assert torch.equal(p.detach(), ref[k])
            assert fb.flat.data_ptr() <= p.data_ptr() < fb.flat.data_ptr() + 4 * fb.n
            assert p.grad is not None and fb.grad.data_ptr() <= p.grad.data_ptr() < fb.grad.data_ptr() + 4 * fb.n
    # build_optimizer's groups (trainer.py:638-646): biases, normalisation weights, the rest
    n_bias = sum(p.numel() for k, p in m.named_parameters() if "bias" in k and p.requires_grad)
    assert int((fb.group == 2).sum()) == n_bias
    assert int((fb.group == 1).sum()) > 0 and int((fb.group == 0).sum()) > 0
    sum((t ** 2).sum() for t in m(torch.rand(2, 3, 64, 64))).backward()
    assert float(fb.grad.abs().sum()) > 0        # autograd accumulated straight into the flat buffer


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mgdt_yolo_b200.modules import Conv
    from mgdt_yolo_b200.train import FlatBucket
    torch.manual_seed(0)
    m = Conv(4, 8, 3).train()                                 # identical replicas
    fb = FlatBucket(m)
    x = torch.full((2, 4, 6, 6), float(rank + 1))
    (m(x) ** 2).sum().backward()
    local = fb.grad.clone()
    scale = fb.all_reduce()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    q.put((rank, scale, bool(torch.allclose(fb.grad, sum(gathered), rtol=1e-6, atol=1e-7))))
    dist.destroy_process_group()


def test_flat_bucket_all_reduce_gloo():
    """One summing all-reduce of the flat gradient bucket == the sum of the ranks' gradients (loss * world_size with
    DDP's averaging, trainer.py:225, 337-338)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert all(ok and scale == 1.0 for _, scale, ok in res), res
