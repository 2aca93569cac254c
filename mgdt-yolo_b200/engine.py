"""Whole-pipeline executor: preprocess -> DetectionModel forward -> decode -> batched NMS.

The reference runs this path as a Python loop over layers plus a Python loop over images with a
host sync per image.  Here the model forward, decode and NMS of one batch are
recorded ONCE into a CUDA graph over a private memory pool (all shapes are static for a given
batch size), so a step is: one preprocess launch + one graph replay, no host work in between and
no device->host traffic except the packed detections.

One Engine drives one GPU.  Multi-GPU inference shards the batch over one process per GPU with no
collective (images are independent).
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import lib
from .modules.base import _version

__all__ = ("Engine",)


class _Slot:
    """One pipeline: device input, graph, outputs, pinned host mirrors, stream."""
    pass


class _Ticket:
    """One submitted batch: the slot that computes it and which of the slot's two input / host-result buffers it uses.
    Quacks like the slot for code that reads `.out`, `.counts`, `.done`, `.stream` (parallel.ShardedEngine)."""

    def __init__(self, slot, k):
        self.slot, self.k = slot, k
        self.done = slot.dones[k]
        self.host_out, self.host_counts = slot.host_outs[k], slot.host_countss[k]

    def __getattr__(self, name):
        return getattr(self.slot, name)


class Engine:
    def __init__(self, model, batch: int, imgsz=640, dtype=torch.bfloat16, device=None, conf=0.25, iou=0.7,
                 max_det=300, multi_label=False, agnostic=False, classes=None, max_nms=30000, max_wh=7680.0,
                 input_dtype=torch.uint8, slots: int = 2, use_graph: bool = True, copy_split: int = 1):
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise RuntimeError("Engine: a CUDA (B200) device is required; there is no CPU fallback")
        h, w = (imgsz, imgsz) if isinstance(imgsz, int) else imgsz
        self.model = model.to(self.device).eval()
        self.batch, self.h, self.w, self.dtype = batch, h, w, dtype
        self.nms_args = dict(conf_thres=conf, iou_thres=iou, multi_label=multi_label, agnostic=agnostic,
                             max_det=max_det, max_nms=max_nms, max_wh=max_wh)
        self.max_det = max_det
        self.input_dtype = input_dtype
        self.ch = self.model.yaml.get("ch", 3)
        self.use_graph = use_graph
        self.launches_per_step = 0
        # bf16: preprocessing is fused into the stem conv (the graph then starts at the raw source buffer)
        self.fused_stem = dtype == torch.bfloat16 and self.model.stem_fusable()
        with torch.cuda.device(self.device):
            self._classes = None if classes is None else torch.as_tensor(list(classes), dtype=torch.int32,
                                                                         device=self.device)
            self.slots = [self._make_slot() for _ in range(max(1, slots))]
        self._next = 0
        # Host -> device input copies: one cudaMemcpyAsync moves ~16.5 GB/s on this platform, three concurrent ones ~29 GB/s
        # (tools/dbg/h2d_bw.py), i.e. 1.37 ms per 39 MB batch of 32 -- as long as the compute.  `copy_split` > 1 splits a
        # batch over that many copy streams; with three slots in flight the slots' own copies already overlap and the split
        # measured slightly slower (e2e 19.8k -> 19.1k images/s), so it is off by default: the end-to-end rate sits at
        # ~85 % of the platform's pinned host -> device ceiling (23.4k images/s of 1.2 MB each).
        self.copy_split = max(1, min(int(copy_split), batch))
        with torch.cuda.device(self.device):
            self._copy_streams = [torch.cuda.Stream(device=self.device) for _ in range(self.copy_split)] if self.copy_split > 1 else []
        # The captured graphs have the packed-weight pointers baked in: keep every pack alive for the engine's lifetime
        # (a later repack by another dtype / a weight update must not free them) and remember the parameter versions
        # the capture saw, so that a replay over changed weights raises instead of silently using stale ones.
        self._pinned = [m.__dict__["_pk"].copy() for m in self.model.modules() if "_pk" in m.__dict__]
        self._pinned.append((dict(ops._STAT_ARENA), dict(ops._TICKETS)))
        self._wstate = [(t, t.data_ptr(), _version(t)) for t in list(self.model.parameters()) + list(self.model.buffers())]

    # ------------------------------------------------------------------ construction
    def _body(self, s):
        """Everything after the input copy, on the current stream, into the slot's static outputs."""
        if self.fused_stem:
            y, _ = self.model.predict_from(s.x0, 1)    # layer 0 ran outside the graph, straight from the source
        else:
            y, _ = self.model(s.x)
        s.pred = y
        nb = lib().mgdt_nms_ws_bytes(self.batch, y.shape[1] - 4, y.shape[2], 1 if self.nms_args["multi_label"] else 0,
                                     self.nms_args["max_nms"])
        if s.ws is None:
            s.ws = torch.empty((nb,), dtype=torch.uint8, device=self.device)
        a = self.nms_args
        ops.nms_packed(y, a["conf_thres"], a["iou_thres"], a["multi_label"], a["agnostic"], a["max_det"], a["max_nms"],
                       a["max_wh"], classes=self._classes, out=s.out, counts=s.counts, ws=s.ws)

    def _make_slot(self):
        s = _Slot()
        B, C, H, W = self.batch, self.ch, self.h, self.w
        s.stream = torch.cuda.Stream(device=self.device)
        s.src = torch.zeros((B, C, H, W), dtype=self.input_dtype, device=self.device)
        s.x = ops.new_act(B, C, H, W, self.dtype, self.device) if not self.fused_stem else None
        s.x0 = None
        # packed detections and per-image counts share ONE allocation (counts = the last B words, viewed as int32): one
        # device -> host copy per batch, one gather per batch in the sharded API
        n_out = B * self.max_det * 6
        s.blob = torch.zeros((n_out + B,), dtype=torch.float32, device=self.device)
        s.out = s.blob[:n_out].view(B, self.max_det, 6)
        s.counts = s.blob[n_out:].view(torch.int32)
        s.ws = None
        # Two submissions may be in flight per slot: while batch i computes, the input of the slot's NEXT batch is already
        # being copied into the other input buffer on the slot's copy stream (the source buffer is only read by the stem,
        # the first launch), and the packed results of batch i sit in their own pinned mirror until collected.
        s.srcs = [s.src, torch.zeros_like(s.src)]
        s.src_free = [None, None]                      # event: the stem that last read srcs[k] has run
        s.copy_stream = torch.cuda.Stream(device=self.device)
        s.host_blobs = [torch.empty((n_out + B,), dtype=torch.float32).pin_memory() for _ in range(2)]
        s.host_outs = [h[:n_out].view(B, self.max_det, 6) for h in s.host_blobs]
        s.host_countss = [h[n_out:].view(torch.int32) for h in s.host_blobs]
        s.dones = [torch.cuda.Event(), torch.cuda.Event()]
        s.outstanding = [False, False]                 # ticket k submitted and not collected yet
        s.turn = 0
        s.host_out, s.host_counts, s.done = s.host_outs[0], s.host_countss[0], s.dones[0]
        s.graph = None
        with torch.no_grad():
            # warm-up (packs weights, sizes the NMS workspace) on the slot's stream, then capture
            s.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s.stream):
                self._head(s, s.src)
                for _ in range(2):
                    self._body(s)
            s.stream.synchronize()
            if self.use_graph:
                c0 = lib().mgdt_launch_count()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s.stream):
                    self._body(s)
                s.graph = g
                self.launches_per_step = int(lib().mgdt_launch_count() - c0) + 1
            else:
                c0 = lib().mgdt_launch_count()
                with torch.cuda.stream(s.stream):
                    self._body(s)
                self.launches_per_step = int(lib().mgdt_launch_count() - c0) + 1
                s.stream.synchronize()
        return s

    # ------------------------------------------------------------------ execution
    def _head(self, s, src):
        """The one launch outside the graph (its source pointer changes per call): preprocessing, fused
        into the stem conv in bf16."""
        if self.fused_stem:
            m0 = self.model.model[0]
            if s.x0 is None:
                s.x0 = m0.forward_image(src)
            else:
                m0.forward_image(src, out=s.x0)
        else:
            ops.preprocess(src, self.dtype, out=s.x)

    def weights_changed(self) -> bool:
        """True if a parameter / buffer was replaced or updated in place since the graphs were captured."""
        return any(t.data_ptr() != p or _version(t) != v for t, p, v in self._wstate)

    def _order_after_producer(self, s, src):
        """The slot's private stream must not read `src` before the stream that produced it is done with it (a device
        tensor the caller has just written on the current stream), and the allocator must not recycle `src` while the
        slot still reads it."""
        if src.is_cuda:
            s.stream.wait_stream(torch.cuda.current_stream(self.device))
            if src.data_ptr() != s.src.data_ptr():
                src.record_stream(s.stream)

    def _run(self, s, src, head_done=None):
        """Enqueue preprocess/stem(src) + forward + decode + NMS on the slot's stream (no sync).  `head_done`: event to
        record once the first launch (the only reader of `src`) is enqueued."""
        self._order_after_producer(s, src)
        with torch.cuda.stream(s.stream), torch.no_grad():
            if not src.is_contiguous():
                src = src.contiguous()      # on the slot's stream, after the wait above
            self._head(s, src)
            if head_done is not None:
                head_done.record(s.stream)
            if s.graph is not None:
                s.graph.replay()
            else:
                self._body(s)

    def step_device(self, src: torch.Tensor, slot: int = 0):
        """Inputs already resident in HBM: `src` is a (B,C,H,W) uint8/float32 device tensor (ordered after the
        caller's current stream).  Returns the slot (outputs in slot.out / slot.counts, on the device, not
        synchronised).  The raw path: it does not re-check the weights (see weights_changed())."""
        s = self.slots[slot]
        self._run(s, src)
        return s

    def submit(self, host_src: torch.Tensor):
        """End-to-end: pinned host batch -> H2D -> pipeline -> D2H of the packed detections.
        Asynchronous; returns a ticket, call `collect(ticket)` for the result.  Consecutive submits rotate over the slots
        so the copies of one batch overlap the kernels of the others, and every slot takes TWO submissions in flight
        (2 x slots in total): the input of a slot's next batch is copied while its current batch computes."""
        if self.use_graph and self.weights_changed():
            raise RuntimeError("Engine: model weights changed after the CUDA graphs were captured; build a new Engine")
        s = self.slots[self._next]
        k = s.turn & 1
        if s.outstanding[k]:
            raise RuntimeError("Engine.submit: more than two uncollected batches on one slot; collect() the oldest first")
        self._next = (self._next + 1) % len(self.slots)
        s.turn += 1
        s.outstanding[k] = True
        src = s.srcs[k]
        if host_src.is_cuda:
            self._order_after_producer(s, host_src)
            with torch.cuda.stream(s.stream):
                src.copy_(host_src, non_blocking=True)
        else:
            cs = s.copy_stream
            if s.src_free[k] is not None:
                cs.wait_event(s.src_free[k])           # the stem of the batch that last used this buffer has read it
            with torch.cuda.stream(cs):
                if self._copy_streams and host_src.is_pinned():
                    ev0 = torch.cuda.Event()
                    ev0.record(cs)
                    for c2, dst, part in zip(self._copy_streams, src.chunk(self.copy_split), host_src.chunk(self.copy_split)):
                        c2.wait_event(ev0)
                        with torch.cuda.stream(c2):
                            dst.copy_(part, non_blocking=True)
                        cs.wait_stream(c2)
                else:
                    src.copy_(host_src, non_blocking=True)
            s.stream.wait_stream(cs)
        if s.src_free[k] is None:
            s.src_free[k] = torch.cuda.Event()
        self._run(s, src, head_done=s.src_free[k])
        with torch.cuda.stream(s.stream):
            s.host_blobs[k].copy_(s.blob, non_blocking=True)
            s.dones[k].record(s.stream)
        return _Ticket(s, k)

    def submit_images(self, ims, auto=False, stride=32):
        """The reference's predictor path for a list of (h, w, 3) BGR uint8 images (numpy or torch, host or device):
        device LetterBox + BGR->RGB + HWC->CHW (preprocess.py) -> pipeline -> scale_boxes back to each original image
        (postprocess.py) -> D2H.  Asynchronous like submit(); collect(slot) returns boxes in original-image pixels."""
        from .postprocess import scale_boxes_params
        from .preprocess import preprocess_images
        if len(ims) != self.batch:
            raise ValueError(f"Engine was built for batches of {self.batch} images, got {len(ims)}")
        if self.use_graph and self.weights_changed():
            raise RuntimeError("Engine: model weights changed after the CUDA graphs were captured; build a new Engine")
        s = self.slots[self._next]
        k = s.turn & 1
        if s.outstanding[k]:
            raise RuntimeError("Engine.submit_images: more than two uncollected batches on one slot; collect() the oldest first")
        self._next = (self._next + 1) % len(self.slots)
        s.turn += 1
        s.outstanding[k] = True
        s.stream.wait_stream(torch.cuda.current_stream(self.device))   # device images written on the caller's stream
        for im in ims:
            if isinstance(im, torch.Tensor) and im.is_cuda:
                im.record_stream(s.stream)
        with torch.cuda.stream(s.stream), torch.no_grad():
            _, metas = preprocess_images(ims, (self.h, self.w), auto=auto, stride=stride, device=self.device, out=s.src)
            prm = torch.tensor([scale_boxes_params((self.h, self.w), m[0]) for m in metas], dtype=torch.float32)
            s.scale_prm = prm.to(self.device, non_blocking=True)
        self._run(s, s.src)
        with torch.cuda.stream(s.stream):
            ops.scale_boxes_packed(s.out, s.counts, s.scale_prm)
            s.host_blobs[k].copy_(s.blob, non_blocking=True)
            s.dones[k].record(s.stream)
        return _Ticket(s, k)

    def collect(self, t):
        """Wait for a submitted batch (its ticket); returns the reference's format: list of (n_i, 6) CPU tensors."""
        if not isinstance(t, _Ticket):                 # a bare slot: its most recent submission
            t = _Ticket(t, (t.turn - 1) & 1)
        t.done.synchronize()
        cnt = t.host_counts.tolist()
        out = t.host_out.clone()                       # one copy out of the pinned mirror, then views per image
        t.slot.outstanding[t.k] = False
        return [out[i, :cnt[i]] for i in range(self.batch)]

    def predict(self, ims, names=None, paths=None, auto=False, stride=32):
        """DetectionPredictor's product for a list of BGR uint8 images (yolo/v8/detect/predict.py:12-30): one Results
        per image with boxes in original-image pixels (submit_images -> collect -> results.build_results)."""
        from .results import build_results
        dets = self.collect(self.submit_images(ims, auto=auto, stride=stride))
        names = names if names is not None else getattr(self.model, "names", None) or {i: str(i) for i in range(self.model.model[-1].nc)}
        return build_results(dets, list(ims), paths, names)

    def __call__(self, images: torch.Tensor):
        """Convenience: images (B,C,H,W) uint8 or float on host or device -> list of (n_i,6) tensors."""
        if images.shape != self.slots[0].src.shape:
            raise ValueError(f"Engine was built for input {tuple(self.slots[0].src.shape)}, got {tuple(images.shape)}")
        if images.dtype != self.input_dtype:
            raise TypeError(f"Engine was built for {self.input_dtype} input")
        if self.use_graph and self.weights_changed():
            raise RuntimeError("Engine: model weights changed after the CUDA graphs were captured; build a new Engine")
        if images.is_cuda:
            s = self.step_device(images)
            s.stream.synchronize()
            cnt = s.counts.tolist()
            return [s.out[i, :cnt[i]].clone() for i in range(self.batch)]
        return self.collect(self.submit(images))
