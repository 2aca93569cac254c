"""CPU, build container only: the plugin rebinding makes the LIVE reference build B200 modules
(skipped on the GPU box, where /root/reference does not exist)."""
import os

import pytest
import torch

from oracle import ref_live

pytestmark = pytest.mark.skipif(not ref_live.available(), reason="live reference not present")


def test_install_rebinds_parse_model_and_facade():
    import mgdt_yolo_b200.modules as M
    from mgdt_yolo_b200 import plugin, tasks
    ref_live.load()
    import ultralytics.nn.tasks as rt
    import ultralytics.yolo.engine.model  # noqa: F401
    ref_model = ref_live.build_model("mspa_c2f_gd_tood_yolov8n.yaml")
    saved = plugin.install()
    try:
        assert rt.MSPA_C2f is M.MSPA_C2f and rt.DetectionModel is tasks.DetectionModel
        # the reference's own parse_model + YAML now builds B200 modules with the reference's state_dict
        yaml_path = os.path.join(ref_live.REFERENCE_DIR, "models", "v8", "mspa_c2f_gd_tood_yolov8n.yaml")
        d = rt.yaml_model_load(yaml_path)
        seq, save = rt.parse_model(d, ch=3, verbose=False)
        assert isinstance(seq[2], M.MSPA_C2f) and isinstance(seq[-1], M.TOODHead)
        ours = rt.DetectionModel(yaml_path, verbose=False)
        assert list(ours.state_dict().keys()) == list(ref_model.state_dict().keys())
        assert all(a.shape == b.shape for a, b in zip(ours.state_dict().values(), ref_model.state_dict().values()))
        ours.load_state_dict(ref_model.state_dict())        # checkpoints transfer unchanged
        assert torch.equal(ours.stride, ref_model.stride) and ours.save == ref_model.save
        from ultralytics import YOLO
        y = YOLO(yaml_path)                                   # the facade (yolo/engine/model.py:104-138)
        assert isinstance(y.model, tasks.DetectionModel)
        import ultralytics.yolo.utils.ops as rops
        from mgdt_yolo_b200.postprocess import non_max_suppression
        assert rops.non_max_suppression is non_max_suppression
    finally:
        plugin.uninstall(saved)
    assert rt.MSPA_C2f is not M.MSPA_C2f
