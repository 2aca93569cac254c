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


def test_update_metrics_batch():
    """update_metrics on the packed NMS output == the reference's per-image loop restated with the oracle
    (scale_boxes on predictions and labels, xywh2xyxy * whwh, process_batch), bit-exact."""
    from mgdt_yolo_b200.postprocess import update_metrics
    from oracle.cases import synth_boxes
    imgsz = (384, 640)
    ori = [(480, 640), (375, 500), (720, 1280), (300, 400)]
    n, max_det = len(ori), 60
    g = torch.Generator().manual_seed(11)
    ratio_pad, dets, det_n, lab_rows, bidx = [], torch.zeros(n, max_det, 6), [60, 17, 0, 9], [], []
    for si, s0 in enumerate(ori):
        gain = min(imgsz[0] / s0[0], imgsz[1] / s0[1])
        pad = (round((imgsz[1] - s0[1] * gain) / 2 - 0.1), round((imgsz[0] - s0[0] * gain) / 2 - 0.1))
        ratio_pad.append(((gain, gain), (float(pad[0]), float(pad[1]))))
        nl = [7, 3, 4, 0][si]
        c = torch.rand(nl, 2, generator=g) * 0.6 + 0.2
        wh = torch.rand(nl, 2, generator=g) * 0.3 + 0.05
        lab_rows.append(torch.cat([torch.randint(0, 2, (nl, 1), generator=g).float(), c, wh], 1))
        bidx.append(torch.full((nl,), float(si)))
        # detections: jittered copies of the labels (letterboxed pixels) + random boxes
        if det_n[si]:
            d = synth_boxes(det_n[si], imgsz, 30 + si)
            if nl:
                lb = lab_rows[-1]
                px = torch.cat([lb[:, 1:3] - lb[:, 3:5] / 2, lb[:, 1:3] + lb[:, 3:5] / 2], 1) * torch.tensor([imgsz[1], imgsz[0]] * 2)
                k = min(nl * 3, det_n[si])
                d[:k] = px.repeat(3, 1)[:k] + (torch.rand(k, 4, generator=g) - 0.5) * 12
            conf = torch.rand(det_n[si], generator=g).sort(descending=True).values
            dets[si, :det_n[si]] = torch.cat([d, conf[:, None], torch.randint(0, 2, (det_n[si], 1), generator=g).float()], 1)
    lab = torch.cat(lab_rows)
    batch = dict(img=torch.empty(n, 3, *imgsz), batch_idx=torch.cat(bidx), cls=lab[:, :1], bboxes=lab[:, 1:], ori_shape=ori, ratio_pad=ratio_pad)
    counts = torch.tensor(det_n, dtype=torch.int32)
    stats, predn = update_metrics(dets.cuda(), counts.cuda(), batch)
    k = 0
    for si in range(n):
        npr, lb = det_n[si], lab_rows[si]
        if npr == 0 and lb.shape[0] == 0:
            continue
        correct, conf, pcls, tcls = stats[k]
        k += 1
        assert torch.equal(tcls.cpu(), lb[:, 0])
        if npr == 0:
            assert correct.shape == (0, 10)
            continue
        pn = O.scale_boxes(imgsz, dets[si, :npr].clone(), ori[si], ratio_pad=ratio_pad[si])
        assert torch.equal(predn[si, :npr].cpu(), pn)
        if lb.shape[0]:
            tbox = O.xywh2xyxy(lb[:, 1:]) * torch.tensor((imgsz[1], imgsz[0], imgsz[1], imgsz[0]), dtype=torch.float32)
            O.scale_boxes(imgsz, tbox, ori[si], ratio_pad=ratio_pad[si])
            ref = O.process_batch(pn, torch.cat((lb[:, :1], tbox), 1))
        else:
            ref = torch.zeros(npr, 10, dtype=torch.bool)
        assert torch.equal(correct.cpu(), ref), f"image {si}"
        assert torch.equal(conf.cpu(), dets[si, :npr, 4]) and torch.equal(pcls.cpu(), dets[si, :npr, 5])
    assert k == len(stats)
