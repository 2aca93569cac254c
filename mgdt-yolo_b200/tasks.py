"""Model graph builder + runner -- mirror of nn/tasks.py (reference): yaml_model_load,
parse_model, BaseModel._predict_once / fuse and DetectionModel, for the detection forward path.

Differences from the reference that do not change results:
  * strides are derived from the graph's spatial bookkeeping instead of a train-mode forward on
    zeros(1,ch,640,640) (tasks.py:243-244) -- no arithmetic ever runs on the CPU;
  * `nn.Upsample` immediately consumed by a `Concat` is not materialised (Concat resamples into
    its slice);
  * `fuse()` is optional: un-fused modules already fold BN at pack time.
"""
from __future__ import annotations

import ast
import contextlib
import math
import os
import re
from copy import deepcopy
from fractions import Fraction

import torch
import torch.nn as nn
import yaml

from . import modules as M
from . import ops
from .modules import (IFM, SPPF, Bottleneck, C2f, Concat, Conv, Detect, InjectionMultiSum_Auto_pool, MSPA_C2f,
                      SimFusion_3in, SimFusion_4in, TOODHead, Upsample)

CFG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cfg")

__all__ = ("DetectionModel", "BaseModel", "parse_model", "yaml_model_load", "guess_model_scale", "make_divisible")


def make_divisible(x, divisor):
    """yolo/utils/torch_utils.py make_divisible (used at tasks.py:640)."""
    return math.ceil(x / divisor) * divisor


def guess_model_scale(model_path):
    """tasks.py:720-735: the n/s/m/l/x letter after 'yolov<d>'."""
    with contextlib.suppress(AttributeError):
        return re.search(r"yolov\d+([nslmx])", os.path.basename(str(model_path))).group(1)
    return ""


def yaml_model_load(path):
    """tasks.py:702-717: '...yolov8n.yaml' -> '...yolov8.yaml' with d['scale'] = 'n'.  Bare file
    names are looked up in this package's cfg/ directory."""
    path = str(path)
    unified = re.sub(r"(\d+)([nslmx])(.+)?$", r"\1\3", path)
    for cand in (unified, path, os.path.join(CFG_DIR, os.path.basename(unified)),
                 os.path.join(CFG_DIR, os.path.basename(path))):
        if os.path.isfile(cand):
            with open(cand, errors="ignore", encoding="utf-8") as f:
                d = yaml.safe_load(f)
            d["scale"] = guess_model_scale(path)
            d["yaml_file"] = path
            return d
    raise FileNotFoundError(f"model yaml '{path}' not found (also looked in {CFG_DIR})")


_NAMES = {k: getattr(M, k) for k in M.__all__}


def parse_model(d, ch, verbose=True):
    """tasks.py:604-699 for the module set of the detection hot path.  Returns (nn.Sequential, save)."""
    max_channels = float("inf")
    nc, act, scales = (d.get(x) for x in ("nc", "activation", "scales"))
    depth, width = (d.get(x, 1.0) for x in ("depth_multiple", "width_multiple"))
    if scales:
        scale = d.get("scale") or tuple(scales.keys())[0]
        depth, width, max_channels = scales[scale]
    if act:
        Conv.default_act = eval(act)  # noqa: S307  same contract as the reference (tasks.py:621)
    ch = [ch]
    layers, save, c2 = [], [], ch[-1]
    for i, (f, n, m, args) in enumerate(d["backbone"] + d["head"]):
        if m == "nn.Upsample":
            m = Upsample
        elif isinstance(m, str) and m.startswith("nn."):
            m = getattr(nn, m[3:])
        elif m in _NAMES:
            m = _NAMES[m]
        else:
            raise NotImplementedError(f"module '{m}' is not part of the MGDT detection path (SURVEY.md §2: out of scope)")
        args = list(args)
        for j, a in enumerate(args):
            if isinstance(a, str):
                if a == "nc":
                    args[j] = nc
                else:
                    with contextlib.suppress(ValueError, SyntaxError):
                        args[j] = ast.literal_eval(a)
        n = n_ = max(round(n * depth), 1) if n > 1 else n
        if m in (Conv, Bottleneck, SPPF, C2f, MSPA_C2f):
            c1, c2 = ch[f], args[0]
            if c2 != nc:
                c2 = make_divisible(min(c2, max_channels) * width, 8)
            args = [c1, c2, *args[1:]]
            if m in (C2f, MSPA_C2f):
                args.insert(2, n)
                n = 1
        elif m is Concat:
            c2 = sum(ch[x] for x in f)
        elif m in (Detect, TOODHead):
            args.append([ch[x] for x in f])  # TOODHead hidc is not width-scaled (tasks.py:664-665)
        elif m is SimFusion_4in:
            c2 = sum(ch[x] for x in f)
        elif m is SimFusion_3in:
            c2 = args[0]
            if c2 != nc:
                c2 = make_divisible(min(c2, max_channels) * width, 8)
            args = [[ch[f_] for f_ in f], c2]
        elif m is IFM:
            c1 = ch[f]
            c2 = sum(args[0])
            args = [c1, *args]
        elif m is InjectionMultiSum_Auto_pool:
            c1 = ch[f[0]]
            c2 = args[0]
            args = [c1, *args]
        else:
            c2 = ch[f]
        m_ = nn.Sequential(*(m(*args) for _ in range(n))) if n > 1 else m(*args)
        t = str(m)[8:-2].replace("__main__.", "")
        m.np = sum(x.numel() for x in m_.parameters())
        m_.i, m_.f, m_.type = i, f, t
        if verbose:
            print(f"{i:>3}{str(f):>20}{n_:>3}{m.np:10.0f}  {t:<45}{str(args):<30}")
        save.extend(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)
        layers.append(m_)
        if i == 0:
            ch = []
        ch.append(c2)
    return nn.Sequential(*layers), sorted(save)


def _downscale(model):
    """Cumulative down-sampling factor of every layer output, from the graph alone."""
    out = []
    for m in model:
        f = m.f
        src = (out[f] if f != -1 else (out[-1] if out else Fraction(1))) if isinstance(f, int) else \
            [out[j] if j != -1 else out[-1] for j in f]
        if isinstance(m, Conv):
            s = src * m.conv.stride[0]
        elif isinstance(m, Upsample):
            s = src / Fraction(m.scale_factor)
        elif isinstance(m, SimFusion_4in):
            s = src[2]
        elif isinstance(m, SimFusion_3in):
            s = src[1]
        elif isinstance(m, (InjectionMultiSum_Auto_pool, Concat)):
            s = src[0]
        elif isinstance(m, (Detect, TOODHead)):
            s = list(src)
        else:
            s = src
        out.append(s)
    return out


class BaseModel(nn.Module):
    """tasks.py:28-216."""

    def forward(self, x, *args, **kwargs):
        if isinstance(x, dict):          # training / validation-loss call of the trainer (tasks.py:42-44)
            return self.loss(x, *args, **kwargs)
        return self.predict(x, *args, **kwargs)

    def loss(self, batch, preds=None):
        """tasks.py:204-216: criterion(preds or forward(batch['img']), batch)."""
        if not hasattr(self, "criterion"):
            self.criterion = self.init_criterion()
        preds = self.forward(batch["img"]) if preds is None else preds
        return self.criterion(preds, batch)

    def predict(self, x, profile=False, visualize=False, augment=False):
        return self._predict_once(x, profile, visualize)

    def _predict_once(self, x, profile=False, visualize=False):
        """tasks.py:65-87.  `profile`/`visualize` are accepted for signature parity and ignored."""
        if self.training:      # differentiable torch-operator forward (train_forward.py); any device
            return self._walk(x)
        ops.require_cuda(x, "model input")
        with torch.cuda.device(x.device):
            return self._walk(x)

    def _walk(self, x):
        y = []
        for m in self.model:
            if m.f != -1:
                x = y[m.f] if isinstance(m.f, int) else [x if j == -1 else y[j] for j in m.f]
            x = m(x)
            y.append(x if m.i in self.save else None)
        return x

    def stem_fusable(self):
        m0 = self.model[0]
        return isinstance(m0, Conv) and m0.can_fuse_preprocess() and bool(ops.lib().mgdt_has_umma())

    def predict_from(self, x, start):
        """Continue the graph walk of _predict_once from layer `start`, x being layer start-1's output."""
        with torch.cuda.device(x.device):
            y = [None] * start
            if start > 0 and self.model[start - 1].i in self.save:
                y[start - 1] = x
            for m in self.model[start:]:
                if m.f != -1:
                    x = y[m.f] if isinstance(m.f, int) else [x if j == -1 else y[j] for j in m.f]
                x = m(x)
                y.append(x if m.i in self.save else None)
            return x

    def predict_image(self, src):
        """Forward from a raw NCHW uint8 (or float32 in [0,1]) batch: the /255 + NCHW->NHWC + bf16
        preprocessing of BasePredictor.preprocess (predictor.py:115-130) is fused into layer 0."""
        ops.require_cuda(src, "model input")
        if not self.stem_fusable():
            return self._predict_once(ops.preprocess(src.contiguous(), torch.bfloat16))
        with torch.cuda.device(src.device):
            return self.predict_from(self.model[0].forward_image(src.contiguous()), 1)

    def fuse(self, verbose=True):
        """tasks.py:121-146: fold every Conv's BatchNorm into its Conv2d."""
        for m in self.model.modules():
            if isinstance(m, Conv) and hasattr(m, "bn"):
                m.fuse()
        return self

    def is_fused(self, thresh=10):
        bn = tuple(v for k, v in nn.__dict__.items() if "Norm" in k)
        return sum(isinstance(v, bn) for v in self.modules()) < thresh

    def _apply(self, fn):
        """tasks.py:171-188: also move the head's stride/anchors."""
        self = super()._apply(fn)
        m = self.model[-1]
        if isinstance(m, (Detect, TOODHead)):
            m.stride = fn(m.stride)
            m.anchors = fn(m.anchors)
            m.strides = fn(m.strides)
        return self

    def load(self, weights, verbose=True):
        """tasks.py:190-202: shape-intersect state_dict transfer."""
        model = weights["model"] if isinstance(weights, dict) and "model" in weights else weights
        csd = model.float().state_dict() if isinstance(model, nn.Module) else model
        own = self.state_dict()
        csd = {k: v for k, v in csd.items() if k in own and own[k].shape == v.shape}
        self.load_state_dict(csd, strict=False)
        if verbose:
            print(f"Transferred {len(csd)}/{len(own)} items from pretrained weights")


class DetectionModel(BaseModel):
    """tasks.py:222-294."""

    def __init__(self, cfg="yolov8n.yaml", ch=3, nc=None, verbose=True):
        super().__init__()
        self.yaml = cfg if isinstance(cfg, dict) else yaml_model_load(cfg)
        ch = self.yaml["ch"] = self.yaml.get("ch", ch)
        if nc and nc != self.yaml["nc"]:
            self.yaml["nc"] = nc
        self.model, self.save = parse_model(deepcopy(self.yaml), ch=ch, verbose=verbose)
        self.names = {i: f"{i}" for i in range(self.yaml["nc"])}
        self.inplace = self.yaml.get("inplace", True)
        m = self.model[-1]
        if isinstance(m, (Detect, TOODHead)):
            m.inplace = self.inplace
            m.stride = torch.tensor([float(s) for s in _downscale(self.model)[-1]])
            self.stride = m.stride
            m.bias_init()
        else:
            self.stride = torch.Tensor([32])
        # an Upsample whose only consumer is the next Concat is resampled straight into the concat buffer
        for i, layer in enumerate(self.model[:-1]):
            nxt = self.model[i + 1]
            if isinstance(layer, Upsample) and isinstance(nxt, Concat) and i not in self.save and -1 in nxt.f:
                layer.defer = True

    def init_criterion(self):
        """tasks.py:293-294: v8DetectionLoss(self) -- here the fused CUDA criterion (train.py, csrc/train.cu)."""
        from .train import v8DetectionLoss
        return v8DetectionLoss(self)
