"""Tail of row f2 on the host: ap_per_class / DetMetrics / ConfusionMatrix / get_stats against the live reference
(yolo/utils/metrics.py, yolo/v8/detect/val.py:124-131) on synthetic validation statistics."""
import numpy as np
import pytest
import torch

from baseline import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout / baseline/_ref not present")


def _stats(seed, n_img=12, nc=3, max_det=40, max_lab=9):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n_img):
        nd = int(torch.randint(0, max_det, (1,), generator=g))
        nl = int(torch.randint(0, max_lab, (1,), generator=g)) if i != 3 else 0
        conf = torch.rand(nd, generator=g)
        level = torch.randint(0, 12, (nd, 1), generator=g)                       # correct up to IoU level `level`
        correct = torch.arange(10).view(1, 10) < level
        out.append((correct, conf, torch.randint(0, nc, (nd,), generator=g).float(), torch.randint(0, nc, (nl,), generator=g).float()))
    return out


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_get_stats_matches_reference(seed):
    from mgdt_yolo_b200.metrics import ap_per_class, get_stats
    ref_loader.load()
    from ultralytics.yolo.utils import metrics as RM
    stats = _stats(seed)
    cols = [torch.cat(x, 0).cpu().numpy() for x in zip(*stats)]
    want = RM.ap_per_class(*cols, plot=False, names={0: "a", 1: "b", 2: "c"})
    got = ap_per_class(*cols)
    for a, b in zip(got, want):
        assert np.allclose(a, b, rtol=0, atol=1e-12)
    dm = RM.DetMetrics(names={0: "a", 1: "b", 2: "c"})
    dm.process(*cols)
    res, m, nt = get_stats(stats, 3, names={0: "a", 1: "b", 2: "c"})
    assert res.keys() == dm.results_dict.keys()
    for k in res:
        assert abs(float(res[k]) - float(dm.results_dict[k])) <= 1e-12, k
    assert np.allclose(m.maps, dm.maps) and nt.tolist() == np.bincount(cols[-1].astype(int), minlength=3).tolist()


def test_get_stats_empty():
    from mgdt_yolo_b200.metrics import get_stats
    res, m, nt = get_stats([], 2)
    assert res["metrics/mAP50(B)"] == 0.0 and nt.tolist() == [0, 0]


@pytest.mark.parametrize("seed", [0, 1])
def test_confusion_matrix_matches_reference(seed):
    from mgdt_yolo_b200.metrics import ConfusionMatrix
    ref_loader.load()
    from ultralytics.yolo.utils import metrics as RM
    g = torch.Generator().manual_seed(10 + seed)
    ours, ref = ConfusionMatrix(3, conf=0.25, iou_thres=0.45), RM.ConfusionMatrix(3, conf=0.25, iou_thres=0.45)
    for _ in range(6):
        nl, nd = int(torch.randint(1, 8, (1,), generator=g)), int(torch.randint(1, 30, (1,), generator=g))
        xy = torch.rand(nl, 2, generator=g) * 400
        wh = torch.rand(nl, 2, generator=g) * 150 + 20
        labels = torch.cat((torch.randint(0, 3, (nl, 1), generator=g).float(), xy, xy + wh), 1)
        pick = torch.randint(0, nl, (nd,), generator=g)
        boxes = labels[pick, 1:] + torch.randn(nd, 4, generator=g) * 12          # jittered copies of the labels
        dets = torch.cat((boxes, torch.rand(nd, 1, generator=g), torch.randint(0, 3, (nd, 1), generator=g).float()), 1)
        ours.process_batch(dets, labels)
        ref.process_batch(dets, labels)
    ours.process_batch(None, torch.tensor([0.0, 2.0]))
    ref.process_batch(None, torch.tensor([0.0, 2.0]))
    assert np.array_equal(ours.matrix, ref.matrix)
    assert all(np.array_equal(a, b) for a, b in zip(ours.tp_fp(), ref.tp_fp()))
