"""Row f3: the training step around the B200 criterion.

  v8DetectionLoss   same constructor / call signature and return value as yolo/utils/loss.py:108-208; the whole
                    criterion -- DFL-expectation decode, task-aligned assigner with the annealed alpha, BCE + CIoU + DFL
                    and the gradient with respect to the head output -- is ONE C-ABI call (mgdt_v8_loss, csrc/train.cu)
                    inside a torch.autograd.Function.
  FlatBucket        parameters, gradients, momentum and the EMA copy of a model as four flat fp32 buffers (the
                    parameters / .grad of the modules become views), so that
                      * the DDP gradient reduction of yolo/engine/trainer.py:225,337-343 is ONE all-reduce over NVLink
                        (parameters that received no gradient -- TOODHead.scale, both reduction_conv biases, SURVEY §3.4 --
                        simply contribute their zeros),
                      * clip_grad_norm_(10) + SGD(nesterov) with the three parameter groups of build_optimizer
                        (trainer.py:614-650, 462-470) is two launches (mgdt_sumsq, mgdt_sgd_step),
                      * ModelEMA.update (yolo/utils/torch_utils.py:347-358) is one launch (mgdt_ema_update) for the
                        parameters plus one torch lerp for the few floating-point buffers (BatchNorm running statistics).
  train_step        forward (train-mode modules, bf16 autocast) -> criterion -> backward -> all-reduce -> optimizer -> EMA.

The model forward / backward themselves run through torch operators (train_forward.py); see DESIGN.md §10.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from ._lib import check, lib

__all__ = ("v8DetectionLoss", "make_anchors", "FlatBucket", "train_step", "synth_targets")


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def make_anchors(feats, strides, grid_cell_offset=0.5):
    """yolo/utils/tal.py:484-500: anchor centres (A, 2) in grid units and the stride of every anchor (A, 1)."""
    pts, sts = [], []
    dtype, device = feats[0].dtype, feats[0].device
    for i, stride in enumerate(strides):
        h, w = feats[i].shape[2:]
        sx = torch.arange(end=w, device=device, dtype=dtype) + grid_cell_offset
        sy = torch.arange(end=h, device=device, dtype=dtype) + grid_cell_offset
        gy, gx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((gx, gy), -1).view(-1, 2))
        sts.append(torch.full((h * w, 1), float(stride), dtype=dtype, device=device))
    return torch.cat(pts), torch.cat(sts)


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, anchors, strides, gt, nc, reg_max, alpha, beta, topk, gains, want):
        b, no, a = pred.shape
        g = gt.shape[1]
        dev = pred.device
        pred = pred.contiguous()
        loss4 = torch.empty(4, dtype=torch.float32, device=dev)
        grad = torch.empty_like(pred)
        ws = torch.empty(lib().mgdt_v8_loss_ws_bytes(b, a, g), dtype=torch.uint8, device=dev)
        outs = [None] * 4
        if want:
            outs = [torch.empty((b, a), dtype=torch.float32, device=dev), torch.empty((b, a), dtype=torch.int32, device=dev),
                    torch.empty((b, a, 4), dtype=torch.float32, device=dev), torch.empty((b, a), dtype=torch.int32, device=dev)]
        ptr = lambda t: None if t is None else t.data_ptr()   # noqa: E731
        check(lib().mgdt_v8_loss(pred.data_ptr(), anchors.data_ptr(), strides.data_ptr(), ptr(gt if g else None), b, a, g, nc, reg_max,
                                 alpha, beta, topk, gains[0], gains[1], gains[2], loss4.data_ptr(), grad.data_ptr(), ptr(outs[0]),
                                 ptr(outs[1]), ptr(outs[2]), ptr(outs[3]), ws.data_ptr(), ws.numel(), _stream(dev)), "mgdt_v8_loss")
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(loss4)
        total = loss4[:3].sum() * b
        return (total, loss4) + tuple(t for t in outs if t is not None)

    @staticmethod
    def backward(ctx, gtotal, *_):
        (grad,) = ctx.saved_tensors
        return (grad * gtotal,) + (None,) * 10


class v8DetectionLoss:
    """Criterion class for computing detection training losses (yolo/utils/loss.py:108-208)."""

    def __init__(self, model):  # model must be de-paralleled
        device = next(model.parameters()).device
        h = getattr(model, "args", None)
        m = model.model[-1]  # Detect() / TOODHead() module
        get = (lambda k, d: (h.get(k, d) if isinstance(h, dict) else getattr(h, k, d))) if h is not None else (lambda k, d: d)
        self.hyp = h
        self.gains = (float(get("box", 7.5)), float(get("cls", 0.5)), float(get("dfl", 1.5)))   # yolo/cfg/default.yaml
        self.stride = m.stride
        self.nc, self.no, self.reg_max = m.nc, m.no, m.reg_max
        self.device = device
        self.epoch = 0                      # call counter (loss.py:127, 205): anneals the assigner's alpha
        self.use_dfl = m.reg_max > 1
        self.topk, self.beta = 10, 8.0      # HeuristicPositiveSampleAssigner_v1(alpha=0.5, beta=8.0) -> TaskAlignedAssigner(topk=10)
        self.last = None                    # assigner outputs of the last call when `keep_assignment` is set (tests)
        self.keep_assignment = False
        if not self.use_dfl:
            raise NotImplementedError("v8DetectionLoss: reg_max = 1 (no DFL) does not occur in this fork")

    def preprocess(self, targets, batch_size, scale_tensor):
        """loss.py:131-148: (n, 6) rows (image, class, xywh normalised) -> (B, max_count, 5) rows (class, xyxy pixels)."""
        if targets.shape[0] == 0:
            return torch.zeros(batch_size, 0, 5, device=self.device)
        img = targets[:, 0].long()
        counts = torch.bincount(img, minlength=batch_size)
        order = torch.argsort(img, stable=True)
        start = torch.cumsum(counts, 0) - counts
        slot = torch.arange(targets.shape[0], device=targets.device) - start[img[order]]
        out = torch.zeros(batch_size, int(counts.max()), 5, device=self.device)
        out[img[order], slot] = targets[order, 1:]
        xywh = out[..., 1:5] * scale_tensor
        half = xywh[..., 2:4] / 2
        out[..., 1:5] = torch.cat((xywh[..., :2] - half, xywh[..., :2] + half), -1)
        return out

    def __call__(self, preds, batch):
        """Sum of the box, cls and dfl losses multiplied by the batch size, and the three detached items."""
        feats = preds[1] if isinstance(preds, tuple) else preds
        b = feats[0].shape[0]
        pred = torch.cat([xi.reshape(b, self.no, -1) for xi in feats], 2).float()
        strides = [float(s) for s in self.stride.tolist()]
        key = tuple(tuple(f.shape[2:]) for f in feats)
        if getattr(self, "_akey", None) != key:
            self._anchors, self._strides = (t.float().contiguous() for t in make_anchors(feats, strides, 0.5))
            self._akey = key
        h, w = feats[0].shape[2:]
        imgsz = torch.tensor([w, h, w, h], device=self.device, dtype=torch.float32) * strides[0]
        targets = torch.cat((batch["batch_idx"].view(-1, 1), batch["cls"].view(-1, 1), batch["bboxes"]), 1).to(self.device).float()
        gt = self.preprocess(targets, b, imgsz).contiguous()
        coff = self.epoch // 161                                            # tal.py:110
        alpha = 0.5 * (100 - coff) / 100                                    # tal.py:266
        res = _LossFn.apply(pred, self._anchors, self._strides, gt, self.nc, self.reg_max, float(alpha), self.beta, self.topk,
                            self.gains, self.keep_assignment)
        total, loss4 = res[0], res[1]
        if self.keep_assignment:
            self.last = dict(target_scores=res[2], target_gt_idx=res[3], target_bboxes=res[4], target_labels=res[5],
                             target_scores_sum=loss4[3], gt=gt)
        self.epoch += 1
        return total, loss4[:3].detach()


def _param_group(module: nn.Module, name: str) -> int:
    """build_optimizer (trainer.py:638-646): 2 = bias (no decay), 1 = normalisation weight (no decay), 0 = weight."""
    if "bias" in name:
        return 2
    if "Norm" in type(module).__name__:
        return 1
    return 0


class FlatBucket:
    """Parameters / gradients / momentum / EMA of a model as flat fp32 buffers + the fused optimizer-side launches."""

    def __init__(self, model: nn.Module, lr=0.01, momentum=0.937, weight_decay=5e-4, nesterov=True, ema_decay=0.9999, ema_tau=2000,
                 max_norm=10.0, process_group=None):
        self.model = model
        seen, params, groups = set(), [], []
        for mname, mod in model.named_modules():
            for pname, p in mod.named_parameters(recurse=False):
                if p.requires_grad and id(p) not in seen:
                    seen.add(id(p))
                    params.append(p)
                    groups.append(_param_group(mod, f"{mname}.{pname}" if mname else pname))
        dev = params[0].device
        self.params = params
        self.n = sum(p.numel() for p in params)
        self.flat = torch.empty(self.n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.mom = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.group = torch.empty(self.n, dtype=torch.uint8, device=dev)
        o = 0
        for p, g in zip(params, groups):
            k = p.numel()
            self.flat[o:o + k].copy_(p.detach().float().reshape(-1))
            p.data = self.flat[o:o + k].view_as(p)
            p.grad = self.grad[o:o + k].view_as(p)
            self.group[o:o + k] = g
            o += k
        self.ema = self.flat.clone()
        self.buffers = [b for b in model.buffers() if b.dtype.is_floating_point]
        self.ema_buffers = [b.detach().clone() for b in self.buffers]
        self.lr = [lr, lr, lr]
        self.wd = [weight_decay, 0.0, 0.0]
        self.momentum, self.nesterov, self.max_norm = momentum, nesterov, max_norm
        self.ema_decay, self.ema_tau, self.updates, self.steps = ema_decay, ema_tau, 0, 0
        self.gnorm_sq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.pg = process_group

    def zero_grad(self):
        self.grad.zero_()
        for p in self.params:          # autograd may have replaced a view (it does not when .grad exists); keep them views
            if p.grad is None or p.grad.data_ptr() < self.grad.data_ptr() or p.grad.data_ptr() >= self.grad.data_ptr() + 4 * self.n:
                raise RuntimeError("FlatBucket: a parameter's .grad no longer aliases the flat gradient buffer")

    def all_reduce(self):
        """The gradient reduction of DDP (trainer.py:225) as ONE all-reduce of the flat bucket.  The trainer multiplies
        the loss by world_size (trainer.py:337-338) and DDP averages the gradients, so the effective gradient is the SUM
        over the ranks: a summing all-reduce, grad_scale 1 (returned for the optimizer launch)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.pg) > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=self.pg)
        return 1.0

    def step(self, grad_scale=1.0):
        """unscale -> clip_grad_norm_(max_norm) -> SGD step -> EMA update (trainer.py:462-470)."""
        dev = self.flat.device
        s = _stream(dev)
        check(lib().mgdt_sumsq(self.grad.data_ptr(), self.n, self.gnorm_sq.data_ptr(), s), "mgdt_sumsq")
        f3 = C.c_float * 3
        check(lib().mgdt_sgd_step(self.flat.data_ptr(), self.grad.data_ptr(), self.mom.data_ptr(), self.group.data_ptr(), self.n,
                                  f3(*self.lr), f3(*self.wd), self.momentum, 1 if self.nesterov else 0, 1 if self.steps == 0 else 0,
                                  self.gnorm_sq.data_ptr(), self.max_norm, float(grad_scale), s), "mgdt_sgd_step")
        self.steps += 1
        self.updates += 1
        d = self.ema_decay * (1 - math.exp(-self.updates / self.ema_tau))
        check(lib().mgdt_ema_update(self.ema.data_ptr(), self.flat.data_ptr(), self.n, d, s), "mgdt_ema_update")
        if self.buffers:
            torch._foreach_mul_(self.ema_buffers, d)
            torch._foreach_add_(self.ema_buffers, [b.detach() for b in self.buffers], alpha=1 - d)

    def ema_state_dict(self):
        """state_dict of the EMA model (ModelEMA.ema.state_dict())."""
        sd = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        names = {id(p): k for k, p in self.model.named_parameters()}
        o = 0
        for p in self.params:
            k = p.numel()
            sd[names[id(p)]] = self.ema[o:o + k].view_as(p).clone()
            o += k
        bnames = {id(b): k for k, b in self.model.named_buffers()}
        for b, e in zip(self.buffers, self.ema_buffers):
            sd[bnames[id(b)]] = e.clone()
        return sd


def synth_targets(batch, per_image=20, nc=2, seed=0, device="cuda"):
    """Synthetic labels of SURVEY §8(d) config #5: ~20 boxes per image, normalised xywh, as the dataloader collates them."""
    g = torch.Generator().manual_seed(seed)
    n = batch * per_image
    cxy = torch.rand(n, 2, generator=g) * 0.8 + 0.1
    wh = torch.rand(n, 2, generator=g) * 0.25 + 0.03
    return {"cls": torch.randint(0, nc, (n, 1), generator=g).float().to(device), "bboxes": torch.cat([cxy, wh], 1).to(device),
            "batch_idx": torch.arange(batch).repeat_interleave(per_image).float().to(device)}


def train_step(model, bucket: FlatBucket, batch, amp_dtype=torch.bfloat16):
    """One iteration of DetectionTrainer._do_train (trainer.py:313-343, 462-470): returns (loss, loss_items)."""
    model.train()
    bucket.zero_grad()
    img = batch["img"]
    if img.dtype == torch.uint8:
        img = img.float() / 255                                            # preprocess_batch (v8/detect/train.py)
    with torch.autocast("cuda", dtype=amp_dtype, enabled=amp_dtype is not None):
        preds = model(img.contiguous(memory_format=torch.channels_last))
    loss, items = model.loss(batch, preds)
    loss.backward()
    scale = bucket.all_reduce()
    bucket.step(scale)
    return loss.detach(), items
