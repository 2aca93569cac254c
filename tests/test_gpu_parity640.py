"""GPU parity at the BENCHMARKED configuration (run on the B200 box: pytest -m gpu): the four BASELINE.json configs
at 640 x 640 -- fp32 validation mode at B = 2 and bf16 at B = 32 -- against the CPU oracle on identical weights
and inputs (every layer output, the decode output, the raw head maps), against the live-reference fixture
tests/golden/model640.npz, and through the CUDA-graph Engine (uint8 in, packed detections out) with the NMS keep set
checked bit-exactly given the engine's own decode tensor.

B = 32 at 640 x 640 exercises code the small fixtures do not: 512-row tiles, paired 16-column units, K-sliced work
items, per-image weight slices for 32 images, 3-D TMA store maps, 148 persistent CTAs, contended fp64 statistics.

Norm (BASELINE.json north_star: "within 1e-2 relative in bf16, 1e-4 in an fp32 validation mode"): fp32 is held to 1e-4
in the max-relative norm max|a - b| / max|b|.  bf16 is held to 1e-2 in the relative L2 norm with the
max-relative tail bounded separately; the derivation, the stated exceptions and the yardstick (the reference's own
modules in bf16 on the same GPU) are in tests/parity.py: bf16_limits.
"""
import pytest
import torch

from tests import parity

pytestmark = pytest.mark.gpu

CFGS = list(parity.BASELINE_CFGS)
FULL = "mspa_c2f_gd_tood_yolov8n.yaml"


@pytest.fixture(scope="module", autouse=True)
def _lib_loaded():
    from mgdt_yolo_b200._lib import lib
    lib()
    assert torch.cuda.is_available()


@pytest.mark.parametrize("cfg", CFGS)
def test_fp32_b2_every_layer(cfg):
    res, y, y_ref = parity.compare_640(cfg, torch.float32, 2)
    for k, (mx, l2) in res.items():
        assert mx <= 1e-4, f"{cfg} {k}: max-rel {mx:.3e}"
    for k, (mx, l2) in parity.check_golden_640(cfg, y).items():     # the live-reference fixture
        assert mx <= 1e-4, f"{cfg} fixture {k}: max-rel {mx:.3e}"


@pytest.mark.parametrize("cfg", CFGS)
def test_bf16_b32_every_layer(cfg):
    res, y, y_ref = parity.compare_640(cfg, torch.bfloat16, 32)
    for k, (mx, l2) in res.items():
        lim_mx, lim_l2 = parity.bf16_limits(k, cfg)
        assert mx <= lim_mx and l2 <= lim_l2, f"{cfg} {k}: max-rel {mx:.3e} (<= {lim_mx}), rel-L2 {l2:.3e} (<= {lim_l2})"


def test_bf16_b2_against_live_reference_fixture():
    res, y, _ = parity.compare_640(FULL, torch.bfloat16, 2)
    for k, (mx, l2) in parity.check_golden_640(FULL, y).items():
        assert l2 <= 1e-2 and mx <= 2e-2, f"fixture {k}: max-rel {mx:.3e}, rel-L2 {l2:.3e}"


def test_bf16_not_worse_than_reference_in_bf16():
    """The yardstick for the bf16 tolerance: the reference's own modules (baseline/_ref) run in bf16 by torch eager on
    this GPU deviate from the fp32 oracle at least as much as the B200 kernels do, layer by layer (relative L2;
    25 % + 1e-3 slack for run-to-run differences of the rounding pattern)."""
    ref = parity.reference_bf16_640(FULL, 8)
    if ref is None:
        pytest.skip("reference copy (baseline/_ref) not present")
    ours, _, _ = parity.compare_640(FULL, torch.bfloat16, 8)
    worse = {k: (ours[k][1], ref[k][1]) for k in ref if k in ours and ours[k][1] > 1.25 * ref[k][1] + 1e-3}
    assert not worse, f"layers where the B200 path deviates more than the reference in bf16 (ours, reference): {worse}"


@pytest.mark.parametrize("dtype,batch", [(torch.bfloat16, 32), (torch.float32, 2)])
def test_engine_640(dtype, batch):
    """DetectionModel through the Engine (CUDA graph, uint8 source, fused stem in bf16): decode output against the
    oracle on the same uint8 images / 255, NMS output == oracle NMS of the engine's own decode tensor (bit-exact)."""
    from mgdt_yolo_b200.engine import Engine
    from oracle import mgdt_oracle as O
    nc = parity.BASELINE_CFGS[FULL]
    m, sd = parity.build_model(FULL, nc=nc, cls_bias=-1.238)
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (batch, 3, 640, 640), dtype=torch.uint8, generator=g)
    eng = Engine(m, batch, 640, dtype, "cuda:0", conf=0.25, iou=0.7, slots=1)
    dets = eng(u8.pin_memory())
    pred = eng.slots[0].pred.float().cpu()
    y_ref, _, _ = parity.oracle_640(FULL, sd, u8.float() / 255, nc, keep_layers=False)
    mx, l2 = parity.errs(pred, y_ref)
    if dtype == torch.bfloat16:
        assert l2 <= 1e-2 and mx <= 2e-2, f"engine decode output vs oracle: max-rel {mx:.3e}, rel-L2 {l2:.3e}"
    else:
        assert mx <= 1e-4, f"engine decode output vs oracle: max-rel {mx:.3e}"
    want = O.non_max_suppression(pred, 0.25, 0.7)
    assert sum(int(t.shape[0]) for t in want) > batch
    for a, b in zip(dets, want):
        assert torch.equal(a.cpu(), b), "NMS keep set differs from the oracle given identical scores and boxes"


def test_engine_two_submissions_in_flight_per_slot():
    """Engine.submit takes 2 x slots uncollected batches (double-buffered inputs and pinned result mirrors); results are
    those of one-at-a-time processing, in ticket order; a further submit without a collect raises."""
    from mgdt_yolo_b200.engine import Engine
    m, _ = parity.build_model(FULL, nc=parity.BASELINE_CFGS[FULL], cls_bias=-1.238)
    eng = Engine(m, 2, 640, torch.bfloat16, "cuda:0", conf=0.25, iou=0.7, slots=2)
    g = torch.Generator().manual_seed(9)
    batches = [torch.randint(0, 256, (2, 3, 640, 640), dtype=torch.uint8, generator=g).pin_memory() for _ in range(4)]
    want = [eng(b) for b in batches]
    tickets = [eng.submit(b) for b in batches]                # 4 = 2 slots x 2
    with pytest.raises(RuntimeError):
        eng.submit(batches[0])
    got = [eng.collect(t) for t in tickets]
    for a, b in zip(got, want):
        assert all(torch.equal(x, y) for x, y in zip(a, b))
    again = eng.collect(eng.submit(batches[1]))
    assert all(torch.equal(x, y) for x, y in zip(again, want[1]))


def test_exported_fused_graph_runs_identically(tmp_path):
    """Row f4 (export of the fused graph): export -> load -> Engine gives the detections of the source model.  The fold
    is computed on the host for the file and on the device at pack time (an fma in the bias on one side only), so the two
    differ in the last bit of some folded values -- a few weights then round to the neighbouring bf16: the decode outputs
    agree within the bf16 tolerance (relative L2 2e-3, max-relative 1e-2), the detections up to threshold crossings."""
    from mgdt_yolo_b200.engine import Engine
    from mgdt_yolo_b200.export import export_fused, load_fused
    m, _ = parity.build_model(FULL, nc=parity.BASELINE_CFGS[FULL], cls_bias=-1.238)
    export_fused(m, str(tmp_path / "full.fused"))
    f = load_fused(str(tmp_path / "full.fused"), "cuda:0")
    g = torch.Generator().manual_seed(21)
    u8 = torch.randint(0, 256, (4, 3, 640, 640), dtype=torch.uint8, generator=g).pin_memory()
    ea = Engine(m, 4, 640, torch.bfloat16, "cuda:0", conf=0.25, iou=0.7, slots=1)
    eb = Engine(f, 4, 640, torch.bfloat16, "cuda:0", conf=0.25, iou=0.7, slots=1)
    a, b = ea(u8), eb(u8)
    assert sum(int(t.shape[0]) for t in a) > 4
    mx, l2 = parity.errs(eb.slots[0].pred, ea.slots[0].pred)
    assert mx <= 1e-2 and l2 <= 2e-3, f"decode output of the exported graph: max-rel {mx:.3e}, rel-L2 {l2:.3e}"
    for x, y in zip(a, b):
        assert abs(x.shape[0] - y.shape[0]) <= max(3, x.shape[0] * 15 // 100)   # random-init heads: many boxes sit at the thresholds
