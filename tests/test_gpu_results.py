"""GPU: Boxes views and the validator's detection <-> label matching are BIT-EXACT with the oracle and with the
fixtures generated from the live reference (tests/golden/results.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import mgdt_oracle as O
from oracle.cases import MATCH_CASES, synth_match

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ci", range(len(MATCH_CASES)), ids=[c[0] for c in MATCH_CASES])
def test_process_batch(ci, golden_dir):
    from mgdt_yolo_b200.postprocess import process_batch
    name, nd, nl, nc, shape = MATCH_CASES[ci]
    dets, labels = synth_match(nd, nl, nc, shape, 500 + ci)
    got = process_batch(dets.cuda(), labels.cuda()).cpu()
    ref = torch.from_numpy(np.load(os.path.join(golden_dir, "results.npz"))[f"pb.{name}"])
    assert got.dtype == torch.bool and torch.equal(got, ref), "differs from the live reference fixture"
    assert torch.equal(got, O.process_batch(dets, labels))


def test_match_batch_packed():
    """Packed batch with per-image counts (an empty image, an image without labels) in one launch."""
    from mgdt_yolo_b200.postprocess import match_batch
    cases = [synth_match(60, 11, 3, (480, 640), 71), synth_match(300, 40, 2, (640, 640), 72), synth_match(5, 4, 1, (64, 64), 73),
             synth_match(20, 6, 2, (320, 320), 74)]
    det_n, lab_n = [60, 300, 0, 20], [11, 40, 4, 0]
    dets = torch.zeros(4, 300, 6)
    labels = torch.zeros(4, 40, 5)
    for i, (d, l) in enumerate(cases):
        dets[i, :d.shape[0]] = d
        labels[i, :l.shape[0]] = l
    got = match_batch(dets.cuda(), torch.tensor(det_n, dtype=torch.int32).cuda(), labels.cuda(),
                      torch.tensor(lab_n, dtype=torch.int32).cuda()).cpu()
    for i, (d, l) in enumerate(cases):
        nd_, nl_ = det_n[i], lab_n[i]
        ref = O.process_batch(d[:nd_], l[:nl_]) if nd_ and nl_ else torch.zeros(nd_, 10, dtype=torch.bool)
        assert torch.equal(got[i, :nd_], ref)
        assert not got[i, nd_:].any()


@pytest.mark.parametrize("ci", range(len(MATCH_CASES)), ids=[c[0] for c in MATCH_CASES])
def test_boxes_views(ci, golden_dir):
    from mgdt_yolo_b200.results import Boxes, Results, build_results
    name, nd, nl, nc, shape = MATCH_CASES[ci]
    dets, _ = synth_match(nd, nl, nc, shape, 500 + ci)
    g = np.load(os.path.join(golden_dir, "results.npz"))
    b = Boxes(dets.cuda(), shape)
    for k in ("xywh", "xyxyn", "xywhn"):
        assert torch.equal(getattr(b, k).cpu(), torch.from_numpy(g[f"{k}.{name}"])), k
    xywh, xyxyn, xywhn = O.boxes_views(dets, shape)
    assert torch.equal(b.xywh.cpu(), xywh) and torch.equal(b.xyxyn.cpu(), xyxyn) and torch.equal(b.xywhn.cpu(), xywhn)
    assert torch.equal(b.conf.cpu(), dets[:, 4]) and torch.equal(b.cls.cpu(), dets[:, 5]) and b.id is None
    r = build_results([dets.cuda()], [np.zeros((*shape, 3), np.uint8)], ["im0.jpg"], {i: f"c{i}" for i in range(nc)})[0]
    assert isinstance(r, Results) and len(r) == nd and r.orig_shape == tuple(shape)
    assert np.array_equal(r.cpu().numpy().boxes.xywh, g[f"xywh.{name}"])       # host copies use the reference's expressions
    assert len(r[:3]) == min(3, nd)
