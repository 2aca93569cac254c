"""Export of the fused graph (row f4; the reference's deployment step is `model.export(format="torchscript")` of the fused
model for pnnx / NCNN, nn/pnnx_conver.py:9, yolo/engine/exporter.py).

The B200 path's deployable artefact is the graph description plus the BN-folded weights: ONE file, loadable without the
training checkpoint or the reference package, from which `load_fused` rebuilds a `DetectionModel` whose convolutions
carry their folded bias (no BatchNorm modules left, `is_fused()` true) -- ready for `Engine`.  The fold is the one
`BaseModel.fuse` performs (`fuse_conv_and_bn`, yolo/utils/torch_utils.py:114-135), in fp32.
"""
from __future__ import annotations

import copy

import torch

FORMAT = "mgdt_b200_fused_v1"

__all__ = ("export_fused", "load_fused", "FORMAT")


def export_fused(model, path: str) -> dict:
    """Write `model` (a DetectionModel, fused or not) as a fused-graph file; returns the saved dictionary's metadata."""
    m = copy.deepcopy(model).float().cpu().eval()
    m.__dict__.pop("criterion", None)
    m.fuse(verbose=False)
    head = m.model[-1]
    blob = {"format": FORMAT, "yaml": copy.deepcopy(m.yaml), "nc": int(m.yaml["nc"]), "names": dict(m.names),
            "stride": [float(s) for s in m.stride.tolist()], "head": type(head).__name__,
            "state_dict": {k: v.detach().clone() for k, v in m.state_dict().items()}}
    torch.save(blob, path)
    return {k: blob[k] for k in ("format", "nc", "names", "stride", "head")}


def load_fused(path: str, device=None):
    """Rebuild the fused DetectionModel from an `export_fused` file (eval mode, on `device` if given)."""
    from .tasks import DetectionModel
    blob = torch.load(path, map_location="cpu", weights_only=False)
    if not isinstance(blob, dict) or blob.get("format") != FORMAT:
        raise ValueError(f"{path}: not a {FORMAT} file")
    m = DetectionModel(blob["yaml"], nc=blob["nc"], verbose=False)
    m.fuse(verbose=False)                                   # same module structure as the exported graph (Conv2d with bias, no BN)
    missing, unexpected = m.load_state_dict(blob["state_dict"], strict=True)
    m.names = dict(blob["names"])
    m.eval()
    return m.to(device) if device is not None else m
