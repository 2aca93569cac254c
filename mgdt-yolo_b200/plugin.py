"""Drop-in registration into the reference package.

The reference resolves module classes by NAME in `ultralytics.nn.tasks`' module globals
(`parse_model`, nn/tasks.py:630) and its engines call `DetectionModel`, `AutoBackend` and
`ops.non_max_suppression`.  `install()` rebinds those names to the B200 implementations, so the
reference's own YAMLs, `YOLO(...)` facade, predictor and validator run unchanged on top of
libmgdt_b200.so.  Nothing here imports the reference unless the caller already has it on sys.path.
"""
from __future__ import annotations

import sys

from . import modules as M
from . import postprocess, tasks

# names the reference looks up (nn/tasks.py:10-14) -> B200 classes
HOT_PATH_NAMES = ("Conv", "Concat", "Bottleneck", "C2f", "MSPA_C2f", "SPPF", "SimFusion_4in", "SimFusion_3in", "IFM",
                  "InjectionMultiSum_Auto_pool", "Detect", "TOODHead")
MODULE_PATHS = {  # pickled checkpoints name classes by module path (trainer.py:411-436)
    "ultralytics.nn.modules.conv": ("Conv", "Concat"),
    "ultralytics.nn.modules.block": ("DFL", "SPPF", "Bottleneck", "C2f", "MSPA_C2f", "SimFusion_4in", "SimFusion_3in",
                                     "IFM", "InjectionMultiSum_Auto_pool", "DyDCNv2", "h_sigmoid"),
    "ultralytics.nn.modules.spr_module": ("SPRModule",),
    "ultralytics.nn.modules.convnextv2": ("ConvNeXtV2_Block",),
    "ultralytics.nn.modules.utils": ("LayerNorm", "GRN"),
    "ultralytics.nn.modules.head": ("Detect", "TOODHead", "TaskDecomposition", "Conv_GN"),
}


_REF_NMS = {}   # the reference's own non_max_suppression, remembered by install() for swap_nms()


def install(ultralytics_pkg=None) -> dict:
    """Rebind the hot-path names inside an already imported reference package.  Returns the dict of
    replaced attributes {qualified name: original object} so `uninstall()` can restore them."""
    if ultralytics_pkg is None:
        ultralytics_pkg = sys.modules.get("ultralytics")
    if ultralytics_pkg is None:
        raise RuntimeError("install(): import the reference package (`ultralytics`) first, or pass it in")
    saved = {}

    def rebind(mod, name, obj):
        saved[f"{mod.__name__}.{name}"] = getattr(mod, name, None)
        setattr(mod, name, obj)

    ref_tasks = sys.modules["ultralytics.nn.tasks"]
    for name in HOT_PATH_NAMES:                       # parse_model's globals() lookup
        rebind(ref_tasks, name, getattr(M, name))
    rebind(ref_tasks, "DetectionModel", tasks.DetectionModel)   # analytic strides: no CPU probe forward
    for path, names in MODULE_PATHS.items():          # checkpoint unpickling + isinstance checks in engines
        mod = sys.modules.get(path)
        if mod is not None:
            for name in names:
                rebind(mod, name, getattr(M, name))
    eng_model = sys.modules.get("ultralytics.yolo.engine.model")
    if eng_model is not None and hasattr(eng_model, "TASK_MAP"):     # YOLO facade: TASK_MAP['detect'][0] (model.py:19-31)
        entry = list(eng_model.TASK_MAP["detect"])
        saved["ultralytics.yolo.engine.model.TASK_MAP.detect"] = tuple(entry)
        entry[0] = tasks.DetectionModel
        eng_model.TASK_MAP["detect"] = entry
    ref_ops = sys.modules.get("ultralytics.yolo.utils.ops")
    if ref_ops is not None:                            # v8/detect/predict.py:14, v8/detect/val.py:65
        _REF_NMS.setdefault("fn", ref_ops.non_max_suppression)
        rebind(ref_ops, "non_max_suppression", postprocess.non_max_suppression)
    return saved


def uninstall(saved: dict):
    tm = saved.pop("ultralytics.yolo.engine.model.TASK_MAP.detect", None)
    if tm is not None:
        sys.modules["ultralytics.yolo.engine.model"].TASK_MAP["detect"] = list(tm)
    for qual, obj in saved.items():
        mod_name, name = qual.rsplit(".", 1)
        mod = sys.modules.get(mod_name)
        if mod is not None and obj is not None:
            setattr(mod, name, obj)


def swap_nms(ours: bool, saved=None):
    """Toggle only the `ops.non_max_suppression` binding of an installed plugin (tests use it to run the untouched
    reference model next to the B200 modules in one process).  Returns the binding that was replaced."""
    ref_ops = sys.modules.get("ultralytics.yolo.utils.ops")
    if ref_ops is None:
        raise RuntimeError("swap_nms(): the reference package is not imported")
    cur = ref_ops.non_max_suppression
    if ours:
        ref_ops.non_max_suppression = postprocess.non_max_suppression
    else:
        ref_ops.non_max_suppression = saved if saved is not None else _REF_NMS.get("fn", cur)
    return cur
