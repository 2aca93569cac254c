"""TEST INFRASTRUCTURE -- the parity case tables shared by oracle/make_golden.py and tests/.

Constructor expressions are evaluated both in the live reference's namespace
(golden generation) and in mgdt_yolo_b200's (parity tests): identical ctor
signatures are part of the drop-in boundary.
"""
MODEL_CFGS = ["yolov8n.yaml", "mspa_c2f_yolov8n.yaml", "mspa_c2f_gd_yolov8n.yaml", "mspa_c2f_gd_tood_yolov8n.yaml",
              "gd_yolov8n.yaml", "gd_thead_yolov8n.yaml", "thead_yolov8n.yaml", "mspa_c2f_thead_yolov8n.yaml"]
LAYER_CFGS = MODEL_CFGS[:4]

# (name, ctor expression evaluated in the reference namespace, input shapes, input-is-list)
MODULE_CASES = [
    ("conv_k3s2", "Conv(16, 32, 3, 2)", [(2, 16, 21, 30)], False),
    ("conv_k1", "Conv(24, 40, 1, 1)", [(2, 24, 9, 13)], False),
    ("conv_k3s1_c8", "Conv(8, 8, 3, 1)", [(1, 8, 17, 19)], False),
    ("bottleneck", "Bottleneck(16, 16, True, 1, ((3, 3), (3, 3)), 1.0)", [(2, 16, 12, 15)], False),
    ("c2f_n2", "C2f(64, 32, 2, True)", [(2, 64, 10, 14)], False),
    ("c2f_n1_noshort", "C2f(48, 64, 1, False)", [(1, 48, 8, 8)], False),
    ("mspa_c2f_n1", "MSPA_C2f(32, 32, 1, True)", [(2, 32, 20, 28)], False),
    ("mspa_c2f_n2", "MSPA_C2f(64, 64, 2, True)", [(2, 64, 11, 13)], False),
    ("sppf", "SPPF(64, 64, 5)", [(2, 64, 7, 9)], False),
    ("simfusion_4in", "SimFusion_4in()", [(2, 8, 32, 48), (2, 16, 16, 24), (2, 24, 8, 12), (2, 32, 4, 6)], True),
    ("simfusion_3in", "SimFusion_3in([16, 32, 32], 32)", [(2, 16, 16, 24), (2, 32, 8, 12), (2, 32, 4, 6)], True),
    ("simfusion_3in_allconv", "SimFusion_3in([8, 16, 24], 32)", [(1, 8, 12, 20), (1, 16, 6, 10), (1, 24, 3, 5)], True),
    ("convnextv2_block", "ConvNeXtV2_Block(96)", [(2, 96, 9, 11)], False),
    ("ifm", "IFM(120, [64, 32])", [(2, 120, 8, 10)], False),
    ("injection", "InjectionMultiSum_Auto_pool(64, 256, [64, 32], 1)", [(2, 64, 12, 16), (2, 96, 6, 8)], True),
    ("injection_flag0", "InjectionMultiSum_Auto_pool(32, 64, [64, 32], 0)", [(1, 32, 10, 14), (1, 96, 5, 7)], True),
    ("conv_gn", "Conv_GN(64, 32, 3)", [(2, 64, 10, 12)], False),
    ("task_decomp", "TaskDecomposition(32, 2, 16)", [(2, 64, 9, 12)], False),
    ("detect", "Detect(5, (32, 64))", [(2, 32, 8, 12), (2, 64, 4, 6)], True),
    ("toodhead", "TOODHead(3, 64, (64,))", [(2, 64, 10, 14)], True),
    ("toodhead_hid128", "TOODHead(80, 128, (128,))", [(1, 128, 5, 6)], True),
]

NMS_CASES = [  # name, nc, anchors, batch, kwargs
    ("nc2_single", 2, 6400, 3, dict(conf_thres=0.25, iou_thres=0.7)),
    ("nc2_multi", 2, 6400, 2, dict(conf_thres=0.001, iou_thres=0.7, multi_label=True)),
    ("nc2_multi_iou45", 2, 3000, 2, dict(conf_thres=0.05, iou_thres=0.45, multi_label=True, max_det=100)),
    ("nc80_single", 80, 8400, 2, dict(conf_thres=0.25, iou_thres=0.7)),
    ("nc80_multi", 80, 1500, 2, dict(conf_thres=0.1, iou_thres=0.6, multi_label=True)),
    ("nc80_agnostic", 80, 2000, 2, dict(conf_thres=0.3, iou_thres=0.5, agnostic=True)),
    ("nc80_classes", 80, 2000, 2, dict(conf_thres=0.2, iou_thres=0.5, classes=[0, 3, 17])),
    ("nc1", 1, 4000, 2, dict(conf_thres=0.4, iou_thres=0.3, multi_label=True)),
    ("nc2_few", 2, 500, 2, dict(conf_thres=0.6, iou_thres=0.7)),
    ("nc2_none", 2, 300, 2, dict(conf_thres=0.9995, iou_thres=0.7)),
]



# pre-processing (LetterBox + BGR->RGB + HWC->CHW): name, source (h, w), new_shape, auto
LETTERBOX_CASES = [
    ("vga_to_384", (480, 640), (384, 384), False),
    ("voc_auto", (375, 500), (320, 320), True),
    ("hd_to_256", (720, 1280), (256, 256), False),
    ("tall_upscale", (100, 37), (320, 320), False),
    ("tiny_rect", (33, 47), (96, 160), False),
    ("odd_auto", (641, 480), (320, 320), True),
    ("same_size", (64, 96), (64, 96), False),
    ("up_to_640", (120, 213), (640, 640), True),
]
# post-processing (scale_boxes + clip_boxes): name, img1 (letterboxed) shape, img0 (original) shape, boxes
SCALE_CASES = [
    ("vga", (640, 640), (480, 640), 300),
    ("voc_auto", (480, 640), (375, 500), 57),
    ("hd", (640, 640), (720, 1280), 300),
    ("tall", (640, 640), (100, 37), 11),
    ("empty", (640, 640), (480, 640), 0),
]


def synth_bgr(h, w, seed):
    import numpy as np
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def synth_boxes(n, shape, seed):
    """xyxy boxes around a letterboxed image of `shape`, some of them outside it (exercise the clip)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    c = torch.rand(n, 2, generator=g) * torch.tensor([shape[1] * 1.2, shape[0] * 1.2]) - torch.tensor([shape[1] * 0.1, shape[0] * 0.1])
    wh = torch.rand(n, 2, generator=g) * 200 + 2
    return torch.cat([c - wh / 2, c + wh / 2], 1)


def module_inputs(name, shapes):
    from mgdt_yolo_b200.synth import synth_images
    return [synth_images(s[0], ch=s[1], h=s[2], w=s[3], seed=100 + i) * 2 - 0.5 for i, s in enumerate(shapes)]


# validation matching (DetectionValidator._process_batch): name, detections, labels, classes, image (h, w)
MATCH_CASES = [
    ("coco_like", 300, 23, 4, (480, 640)),
    ("crowded", 300, 160, 2, (640, 640)),
    ("few", 7, 3, 1, (375, 500)),
    ("one_label", 40, 1, 2, (640, 640)),
    ("no_match", 25, 9, 80, (640, 640)),
    ("single_det", 1, 12, 2, (320, 320)),
]


def synth_match(nd, nl, nc, shape, seed):
    """Labels (cls, xyxy) and detections (xyxy, conf, cls): most detections are jittered copies of labels (several per
    label, at every IoU level between 0.3 and 1), the rest are random boxes; no two labels coincide, so no IoU ties."""
    import torch
    g = torch.Generator().manual_seed(seed)
    h, w = shape
    c = torch.rand(nl, 2, generator=g) * torch.tensor([w * 0.8, h * 0.8]) + torch.tensor([w * 0.1, h * 0.1])
    wh = torch.rand(nl, 2, generator=g) * torch.tensor([w * 0.3, h * 0.3]) + 8
    lab_box = torch.cat([c - wh / 2, c + wh / 2], 1)
    lab_cls = torch.randint(0, nc, (nl, 1), generator=g).float()
    labels = torch.cat([lab_cls, lab_box], 1)
    src = torch.randint(0, nl, (nd,), generator=g)
    jit = (torch.rand(nd, 4, generator=g) - 0.5) * torch.rand(nd, 1, generator=g) * 0.6
    box = lab_box[src] + jit * torch.cat([wh[src], wh[src]], 1)
    rnd = torch.rand(nd, generator=g) < 0.25
    rb = synth_boxes(nd, shape, seed + 1)
    box[rnd] = rb[rnd]
    cls = lab_cls[src, 0].clone()
    flip = torch.rand(nd, generator=g) < 0.15
    cls[flip] = torch.randint(0, nc, (int(flip.sum()),), generator=g).float()
    conf = torch.rand(nd, generator=g)
    dets = torch.cat([box, conf[:, None], cls[:, None]], 1)
    return dets[conf.argsort(descending=True)].contiguous(), labels.contiguous()
