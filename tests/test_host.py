"""CPU tests of the host logic: the C-ABI library loads and exports every declared symbol, the
graph builder reproduces the reference's channel arithmetic / state_dict keys, loud failures."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from mgdt_yolo_b200 import _lib
    from mgdt_yolo_b200.build import build
    build()
    header = open(os.path.join(ROOT, "include", "mgdt_b200.h")).read()
    declared = set(re.findall(r"\b(mgdt_[a-z0-9_]+)\s*\(", header)) - {"mgdt_conv_args", "mgdt_decode_level"}
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name)
    assert L.mgdt_abi_version() == _lib.ABI_VERSION


def test_sass_is_sm100a():
    lib = os.path.join(ROOT, "mgdt-yolo_b200", "libmgdt_b200.so")
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.parametrize("cfg", ["yolov8n.yaml", "mspa_c2f_yolov8n.yaml", "mspa_c2f_gd_yolov8n.yaml",
                                 "mspa_c2f_gd_tood_yolov8n.yaml", "gd_yolov8n.yaml", "gd_thead_yolov8n.yaml",
                                 "thead_yolov8n.yaml", "mspa_c2f_thead_yolov8n.yaml"])
def test_model_state_dict_matches_reference(cfg, golden_dir):
    """Same keys, same shapes, same parameter count as the live reference's DetectionModel."""
    from mgdt_yolo_b200.tasks import DetectionModel
    g = np.load(os.path.join(golden_dir, f"model_{cfg[:-5]}.npz"))
    m = DetectionModel(cfg, verbose=False)
    sd = m.state_dict()
    assert list(sd.keys()) == list(g["keys"])
    assert [",".join(map(str, v.shape)) for v in sd.values()] == list(g["shapes"])
    assert sum(p.numel() for p in m.parameters()) == int(g["n_params"])


def test_strides_and_save_list():
    from mgdt_yolo_b200.tasks import DetectionModel
    m = DetectionModel("mspa_c2f_gd_tood_yolov8n.yaml", verbose=False)
    assert m.stride.tolist() == [8.0] and m.save == [2, 2, 4, 4, 6, 6, 9, 11, 15]
    assert m.model[-1].reg_max == 16 and m.model[-1].no == 2 + 64
    m = DetectionModel("yolov8n.yaml", verbose=False)
    assert m.stride.tolist() == [8.0, 16.0, 32.0] and m.save == [4, 6, 9, 12, 15, 18, 21]
    assert m.model[-1].reg_max == 4  # the fork's Detect (head.py:145)
    s = DetectionModel("yolov8s.yaml", verbose=False)
    assert s.model[0].conv.out_channels == 32


def test_no_cpu_fallback():
    from mgdt_yolo_b200.modules import Conv
    from mgdt_yolo_b200.postprocess import non_max_suppression
    from mgdt_yolo_b200.tasks import DetectionModel
    with pytest.raises(RuntimeError, match="no CPU"):
        Conv(3, 8, 3).eval()(torch.zeros(1, 3, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU"):
        DetectionModel("yolov8n.yaml", verbose=False).eval()(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError):
        non_max_suppression(torch.zeros(1, 6, 10), 0.25, 0.7)
    with pytest.raises(AssertionError):
        non_max_suppression(torch.zeros(1, 6, 10), 1.25, 0.7)


def test_fuse_matches_reference_formula():
    from mgdt_yolo_b200.modules import Conv
    from mgdt_yolo_b200.synth import synth_state_dict
    from oracle import mgdt_oracle as O
    c = Conv(8, 16, 3, 2)
    sd = synth_state_dict(c.state_dict(), seed=3)
    c.load_state_dict(sd)
    folded = O.fold_bn(sd)
    c.fuse()
    assert not hasattr(c, "bn")
    assert torch.allclose(c.conv.weight, folded["conv.weight"], atol=1e-6)
    assert torch.allclose(c.conv.bias, folded["conv.bias"], atol=1e-6)


def test_host_side_queries_of_the_new_entry_points():
    """Pure host queries (no GPU): supported widths / packed sizes of the fused MSPA branch chain and the tcgen05 weight
    images; argument validation that fails before any launch returns a negative code and a message."""
    import ctypes as C
    from mgdt_yolo_b200 import _lib
    L = _lib.lib()
    assert L.mgdt_mspa_front_supported(8, 3) == 1 and L.mgdt_mspa_front_supported(64, 3) == 1
    assert L.mgdt_mspa_front_supported(96, 3) == 0 and L.mgdt_mspa_front_supported(16, 0) == 0
    # [stage][n-tile][k-block][lane]{b0, b1}: 3 stages x (64/8) x (64/16) x 32 lanes x 8 bytes
    assert L.mgdt_mspa_front_packed_bytes(64, 3) == 3 * 8 * 4 * 32 * 8
    assert L.mgdt_mspa_front_packed_bytes(8, 3) == 3 * 1 * 1 * 32 * 8          # K padded to one 16-channel block
    assert L.mgdt_mspa_front_packed_bytes(24, 3) == 0
    assert L.mgdt_conv_umma_packed_bytes(384, 96, 1, 1) == 384 * 96 * 2        # K-sliced, no padding for this shape
    assert L.mgdt_conv_umma_packed_bytes(7, 16, 1, 1) == 0                     # Cin % 8 != 0: not a tcgen05 shape
    assert L.mgdt_box_convert(None, 6, 0, 1, 1.0, 1.0, None, None) == 0        # n = 0 is a no-op
    assert L.mgdt_box_convert(None, 6, 4, 1, 1.0, 1.0, None, None) < 0 and b"null" in L.mgdt_last_error()
    assert L.mgdt_match_batch(None, 6, None, 300, None, None, 8, None, 64, None, 2, None) < 0   # niou > 32
    assert L.mgdt_stats_finish(None, 4, 1, 8, 8, 32, 1, 1, None, None, None, None) < 0


def test_export_fused_round_trip(tmp_path):
    """Row f4: the fused-graph file rebuilds a BN-free DetectionModel with the folded weights of BaseModel.fuse."""
    import torch
    from mgdt_yolo_b200.export import export_fused, load_fused
    from mgdt_yolo_b200.synth import synth_state_dict
    from mgdt_yolo_b200.tasks import DetectionModel
    m = DetectionModel("mspa_c2f_gd_tood_yolov8n.yaml", nc=2, verbose=False)
    m.load_state_dict(synth_state_dict(m.state_dict(), seed=1))
    meta = export_fused(m, str(tmp_path / "m.fused"))
    assert meta["head"] == "TOODHead" and meta["nc"] == 2
    assert not m.is_fused()                                   # the source model is untouched
    f = load_fused(str(tmp_path / "m.fused"))
    assert f.is_fused() and not any(isinstance(x, torch.nn.BatchNorm2d) for x in f.modules())
    ref = DetectionModel("mspa_c2f_gd_tood_yolov8n.yaml", nc=2, verbose=False)
    ref.load_state_dict(synth_state_dict(ref.state_dict(), seed=1))
    ref.fuse(verbose=False)
    a, b = ref.state_dict(), f.state_dict()
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)


def test_bench_clock_sampler_degrades_without_a_gpu():
    """bench.py samples SM clocks through NVML (nvidia-smi as the fallback); with neither it must report None, not raise."""
    import sys
    import time
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    s.start()
    t0 = time.time()
    time.sleep(0.05)
    out = s.stop(t0, time.time())
    assert out is None or {"sm_mhz", "sm_max_mhz", "reasons", "samples", "source"} <= set(out)
