"""TEST / BENCH INFRASTRUCTURE -- loader for the UNMODIFIED reference.

Two homes, tried in this order:
  /root/reference                  the read-only checkout of the build container (golden-vector generation,
                                   `oracle/make_golden.py`, CPU tests that pin the oracle against the live reference)
  <repo>/baseline/_ref/ultralytics the copy `baseline/install_ref.py` makes (git-ignored, NOT gpurun-ignored, so it
                                   travels to the GPU box): `bench.py --impl reference`, the `gpu_eager_baseline`
                                   leg and the plugin tests on the B200 run the reference's own code from there.
Nothing in the product path (mgdt_yolo_b200/) imports this module.

The reference imports itself as `ultralytics.*` (__init__.py:5-10) and needs five
third-party modules that are not installed here; they are
stubbed in `sys.modules` before import:

  timm / timm.models.layers   nn/tasks.py:6, nn/modules/convnextv2.py:11   (trunc_normal_, DropPath)
  mmcv.cnn                    nn/modules/head.py:13, block.py:16            (ConvModule, Scale, build_norm_layer)
  mmcv.ops.modulated_deform_conv  nn/modules/block.py:17                    (ModulatedDeformConv2d -> torchvision.ops.deform_conv2d)
  mmengine.model              nn/modules/head.py:14                         (normal_init)
  matplotlib / seaborn        yolo/utils/__init__.py:19, checks.py:20       (MagicMock)

DCNv2 stand-in: torchvision.ops.deform_conv2d with mmcv's conventions assumed
(offset channels interleaved (dy,dx) per tap; mask per tap).  mmcv itself is
absent, so this boundary is "parity unpinned" (SURVEY.md §8(c)).
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIVE_DIR = "/root/reference"
SHIPPED_DIR = os.path.join(ROOT, "baseline", "_ref", "ultralytics")


def _is_ref(d) -> bool:
    return os.path.isdir(os.path.join(d, "nn", "modules")) and os.path.isfile(os.path.join(d, "nn", "tasks.py"))


REFERENCE_DIR = LIVE_DIR if _is_ref(LIVE_DIR) else SHIPPED_DIR


def available() -> bool:
    return _is_ref(REFERENCE_DIR)


def kind() -> str:
    """'live' (build container checkout), 'shipped' (baseline/_ref copy) or 'absent'."""
    return "absent" if not available() else ("live" if REFERENCE_DIR == LIVE_DIR else "shipped")


def _install_stubs():
    import torch
    import torch.nn as nn
    import torchvision

    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        timm_models = types.ModuleType("timm.models")
        timm_layers = types.ModuleType("timm.models.layers")

        class DropPath(nn.Identity):
            def __init__(self, p=0.0):
                super().__init__()

        timm_layers.trunc_normal_ = nn.init.trunc_normal_
        timm_layers.DropPath = DropPath
        timm.models = timm_models
        timm_models.layers = timm_layers
        sys.modules.update({"timm": timm, "timm.models": timm_models, "timm.models.layers": timm_layers})

    if "mmcv" not in sys.modules:
        mmcv = types.ModuleType("mmcv")
        mmcv_cnn = types.ModuleType("mmcv.cnn")
        mmcv_ops = types.ModuleType("mmcv.ops")
        mmcv_mdc = types.ModuleType("mmcv.ops.modulated_deform_conv")

        class ConvModule(nn.Module):
            def __init__(self, cin, cout, k, stride=1, padding=0, conv_cfg=None, norm_cfg=None, bias=True):
                super().__init__()
                self.conv = nn.Conv2d(cin, cout, k, stride, padding, bias=bias)
                self.activate = nn.ReLU(inplace=True)

            def forward(self, x):
                return self.activate(self.conv(x))

        class Scale(nn.Module):
            def __init__(self, scale=1.0):
                super().__init__()
                self.scale = nn.Parameter(torch.tensor(scale, dtype=torch.float))

            def forward(self, x):
                return x * self.scale

        def build_norm_layer(cfg, num_features):
            assert cfg["type"] == "GN"
            return "gn", nn.GroupNorm(cfg["num_groups"], num_features)

        def build_activation_layer(cfg):
            raise NotImplementedError

        class ModulatedDeformConv2d(nn.Module):
            def __init__(self, cin, cout, k, stride=1, padding=0, dilation=1, groups=1, deform_groups=1, bias=True):
                super().__init__()
                self.stride, self.padding, self.dilation = stride, padding, dilation
                self.weight = nn.Parameter(torch.empty(cout, cin // groups, k, k))
                nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
                self.bias = nn.Parameter(torch.zeros(cout)) if bias else None

            def forward(self, x, offset, mask):
                if x.dtype == torch.bfloat16:   # torchvision has no bf16 deformable_im2col: compute in fp32, round the result
                    b = None if self.bias is None else self.bias.float()
                    return torchvision.ops.deform_conv2d(x.float(), offset.float(), self.weight.float(), b, self.stride,
                                                         self.padding, self.dilation, mask.float()).to(x.dtype)
                return torchvision.ops.deform_conv2d(x, offset, self.weight, self.bias, self.stride, self.padding,
                                                     self.dilation, mask)

        mmcv_cnn.ConvModule, mmcv_cnn.Scale = ConvModule, Scale
        mmcv_cnn.build_norm_layer, mmcv_cnn.build_activation_layer = build_norm_layer, build_activation_layer
        mmcv_mdc.ModulatedDeformConv2d = ModulatedDeformConv2d
        mmcv.cnn, mmcv.ops, mmcv_ops.modulated_deform_conv = mmcv_cnn, mmcv_ops, mmcv_mdc
        sys.modules.update({"mmcv": mmcv, "mmcv.cnn": mmcv_cnn, "mmcv.ops": mmcv_ops,
                            "mmcv.ops.modulated_deform_conv": mmcv_mdc})

    if "mmengine" not in sys.modules:
        mmengine = types.ModuleType("mmengine")
        mmengine_model = types.ModuleType("mmengine.model")

        def normal_init(module, mean=0, std=1, bias=0):
            nn.init.normal_(module.weight, mean, std)
            if getattr(module, "bias", None) is not None:
                nn.init.constant_(module.bias, bias)

        mmengine_model.normal_init = normal_init
        mmengine.model = mmengine_model
        sys.modules.update({"mmengine": mmengine, "mmengine.model": mmengine_model})

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.font_manager", "matplotlib.image", "seaborn"):   # thop is optional in the reference (nn/tasks.py:22-25)
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = mock.MagicMock()


_LOADED = None


def load():
    """Import the live reference as `ultralytics` and return the package."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError("reference not present: neither /root/reference nor baseline/_ref (run baseline/install_ref.py "
                           "in the build container)")
    os.environ.setdefault("YOLO_VERBOSE", "false")
    _install_stubs()
    if REFERENCE_DIR == SHIPPED_DIR:
        link_root = os.path.dirname(SHIPPED_DIR)          # baseline/_ref holds the package directory itself
    else:
        link_root = os.path.join(tempfile.gettempdir(), "mgdt_ref_link")
        os.makedirs(link_root, exist_ok=True)
        link = os.path.join(link_root, "ultralytics")
        if not os.path.islink(link):
            os.symlink(REFERENCE_DIR, link)
    if link_root not in sys.path:
        sys.path.insert(0, link_root)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import ultralytics  # noqa: F401
        import ultralytics.nn.tasks  # noqa: F401
        import ultralytics.yolo.utils.ops  # noqa: F401
    _LOADED = sys.modules["ultralytics"]
    return _LOADED


def build_model(cfg: str, nc=None):
    """DetectionModel(cfg) from the reference's own YAML (models/v8/<cfg>)."""
    load()
    import torch
    from ultralytics.nn.tasks import DetectionModel
    path = os.path.join(REFERENCE_DIR, "models", "v8", cfg)
    torch.manual_seed(0)
    return DetectionModel(path, nc=nc, verbose=False)
