"""GPU: device LetterBox / preprocess and scale_boxes are BIT-EXACT with the oracle and the live-reference fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import mgdt_oracle as O
from oracle.cases import LETTERBOX_CASES, SCALE_CASES, synth_bgr, synth_boxes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ci", range(len(LETTERBOX_CASES)), ids=[c[0] for c in LETTERBOX_CASES])
def test_preprocess_images(ci, golden_dir):
    from mgdt_yolo_b200.preprocess import LetterBox, preprocess_images
    name, shape, new_shape, auto = LETTERBOX_CASES[ci]
    img = synth_bgr(shape[0], shape[1], 300 + ci)
    out, metas = preprocess_images([img, img[::-1].copy()], new_shape, auto=auto, stride=32)
    ref = np.load(os.path.join(golden_dir, "prepost.npz"))[f"lb.{name}"]
    assert np.array_equal(out[0].cpu().numpy(), ref), "differs from the live reference (cv2) fixture"
    assert np.array_equal(out[1].cpu().numpy(), O.preprocess_images([img[::-1].copy()], new_shape, auto=auto)[0])
    assert metas[0][0] == tuple(shape)
    hwc = LetterBox(new_shape, auto=auto, stride=32)(image=img)     # the reference's HWC / BGR view of the same thing
    assert np.array_equal(hwc.cpu().numpy(), O.letterbox(img, new_shape, auto=auto))


def test_preprocess_strided_rows():
    """A source image that is a column crop of a wider buffer (row pitch > 3 * w)."""
    from mgdt_yolo_b200 import ops
    big = torch.from_numpy(synth_bgr(90, 200, 7)).cuda()
    view = big[:, 20:150]                                              # (90, 130, 3), pitch 600
    dst = torch.empty((3, 128, 160), dtype=torch.uint8, device="cuda")
    new_unpad, (top, bottom, left, right), _, _ = O.letterbox_params((90, 130), (128, 160))
    ops.letterbox_u8(view, dst, (new_unpad[1], new_unpad[0]), (top, left))
    ref = O.preprocess_images([view.cpu().numpy()], (128, 160))[0]
    assert np.array_equal(dst.cpu().numpy(), ref)


@pytest.mark.parametrize("ci", range(len(SCALE_CASES)), ids=[c[0] for c in SCALE_CASES])
def test_scale_boxes(ci, golden_dir):
    from mgdt_yolo_b200.postprocess import scale_boxes
    name, s1, s0, n = SCALE_CASES[ci]
    b = synth_boxes(n, s1, 400 + ci)
    out = scale_boxes(s1, b.clone().cuda(), s0)
    ref = torch.from_numpy(np.load(os.path.join(golden_dir, "prepost.npz"))[f"sb.{name}"])
    assert torch.equal(out.cpu(), ref)
    # with extra columns (conf, cls) and an explicit ratio_pad, against the oracle
    full = torch.cat([synth_boxes(n, s1, 9), torch.rand(n, 2)], 1)
    rp = ((0.8125, 0.8125), (12.0, 3.0))
    got = scale_boxes(s1, full.clone().cuda(), s0, ratio_pad=rp)
    assert torch.equal(got.cpu(), O.scale_boxes(s1, full.clone(), s0, ratio_pad=rp))


def test_scale_boxes_packed_batch():
    """Packed NMS output of a batch with per-image counts and per-image original shapes."""
    from mgdt_yolo_b200 import ops
    from mgdt_yolo_b200.postprocess import scale_boxes_params
    shapes0 = [(480, 640), (375, 500), (720, 1280)]
    dets = torch.cat([synth_boxes(40, (640, 640), 50 + i).unsqueeze(0) for i in range(3)])
    dets = torch.cat([dets, torch.rand(3, 40, 2)], 2).contiguous()
    counts = torch.tensor([40, 7, 0], dtype=torch.int32)
    prm = torch.tensor([scale_boxes_params((640, 640), s) for s in shapes0], dtype=torch.float32)
    out = ops.scale_boxes_packed(dets.clone().cuda(), counts.cuda(), prm.cuda()).cpu()
    for i, s0 in enumerate(shapes0):
        c = int(counts[i])
        assert torch.equal(out[i, :c], O.scale_boxes((640, 640), dets[i, :c].clone(), s0))
        assert torch.equal(out[i, c:], dets[i, c:])                   # rows past the count are untouched


def test_engine_submit_images():
    """Engine.submit_images == preprocess_images -> step_device -> scale_boxes, composed by hand."""
    from mgdt_yolo_b200 import ops
    from mgdt_yolo_b200.engine import Engine
    from mgdt_yolo_b200.postprocess import scale_boxes_params
    from mgdt_yolo_b200.preprocess import preprocess_images
    from mgdt_yolo_b200.synth import raise_cls_bias, synth_state_dict
    from mgdt_yolo_b200.tasks import DetectionModel
    model = DetectionModel("mspa_c2f_gd_tood_yolov8n.yaml", nc=2, verbose=False)
    model.load_state_dict(raise_cls_bias(synth_state_dict(model.state_dict(), seed=1), -1.238))
    eng = Engine(model, 2, (128, 160), torch.bfloat16, "cuda:0", conf=0.25, iou=0.7, slots=1)
    ims = [synth_bgr(90, 130, 1), synth_bgr(200, 120, 2)]
    got = eng.collect(eng.submit_images(ims))
    batch, metas = preprocess_images(ims, (128, 160))
    s = eng.step_device(batch)
    s.stream.synchronize()
    prm = torch.tensor([scale_boxes_params((128, 160), m[0]) for m in metas], dtype=torch.float32, device="cuda")
    with torch.cuda.stream(s.stream):
        ops.scale_boxes_packed(s.out, s.counts, prm)
    s.stream.synchronize()
    cnt = s.counts.tolist()
    assert sum(cnt) > 0
    for i in range(2):
        assert torch.equal(got[i], s.out[i, :cnt[i]].cpu())
        assert float(got[i][:, [0, 2]].max()) <= ims[i].shape[1] and float(got[i][:, [1, 3]].max()) <= ims[i].shape[0]
    res = eng.predict(ims, paths=["a.jpg", "b.jpg"])                  # the same thing as Results objects
    for i in range(2):
        assert res[i].orig_shape == ims[i].shape[:2] and torch.equal(res[i].boxes.data, got[i])
        assert torch.equal(res[i].boxes.xywh, O.boxes_views(got[i], ims[i].shape[:2])[0])
