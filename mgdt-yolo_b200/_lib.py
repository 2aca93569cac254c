"""ctypes binding of libmgdt_b200.so (include/mgdt_b200.h).

There is no fallback: if the shared library is missing or a call fails, a
RuntimeError is raised with the library's own message.  `python -m
mgdt_yolo_b200.build` (or `__graft_entry__.build()`) compiles it with nvcc for
sm_100a; the file travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MGDT_LIB") or os.path.join(HERE, "libmgdt_b200.so")   # MGDT_LIB: A/B runs of two builds in one session

ABI_VERSION = 5  # include/mgdt_b200.h MGDT_ABI_VERSION
F32, BF16 = 0, 1
ACT_NONE, ACT_SILU, ACT_RELU, ACT_SIGMOID, ACT_HSIGMOID, ACT_GELU = range(6)
RS_COPY, RS_AVGPOOL, RS_BILINEAR, RS_NEAREST = range(4)

vp, i32, f32, sz = C.c_void_p, C.c_int32, C.c_float, C.c_size_t


class ConvArgs(C.Structure):
    _fields_ = [("x", vp), ("w", vp), ("bias", vp), ("y", vp), ("pre_add", vp), ("in_scale", vp), ("pix_scale", vp),
                ("residual", vp),
                ("N", i32), ("H", i32), ("W", i32), ("Cin", i32), ("Cout", i32),
                ("kh", i32), ("kw", i32), ("stride", i32), ("pad", i32),
                ("x_cs", i32), ("y_cs", i32), ("add_cs", i32), ("ps_cs", i32), ("res_cs", i32),
                ("act", i32), ("in_relu", i32), ("dtype", i32), ("impl", i32), ("w_umma", vp), ("w_umma_f16", i32),
                ("stat_acc", vp), ("stat_q", i32), ("stat_sq", i32), ("stat_copies", i32), ("w_per_image", i32), ("act_cols", i32)]


class StatsFin(C.Structure):
    fp = C.POINTER(C.c_float)
    _fields_ = [("kind", i32), ("p0", vp), ("p1", vp), ("p2", vp), ("p3", vp), ("i0", i32), ("i1", i32), ("i2", i32),
                ("f0", C.c_float), ("o0", vp), ("o1", vp)]


class DecodeLevel(C.Structure):
    _fields_ = [("raw", vp), ("H", i32), ("W", i32), ("cs", i32), ("stride", f32)]


# name -> (restype, argtypes); every symbol include/mgdt_b200.h declares
SIGNATURES = {
    "mgdt_abi_version": (C.c_int, []),
    "mgdt_last_error": (C.c_char_p, []),
    "mgdt_launch_count": (C.c_ulonglong, []),
    "mgdt_has_umma": (C.c_int, []),
    "mgdt_debug_set_trace": (None, [vp]),
    "mgdt_set_pdl": (None, [i32]),
    "mgdt_set_option": (C.c_int, [C.c_char_p, i32]),
    "mgdt_letterbox_u8": (C.c_int, [vp, i32, i32, i32, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    "mgdt_box_convert": (C.c_int, [vp, i32, i32, i32, f32, f32, vp, vp]),
    "mgdt_match_batch": (C.c_int, [vp, i32, vp, i32, vp, vp, i32, vp, i32, vp, i32, vp]),
    "mgdt_scale_boxes": (C.c_int, [vp, i32, vp, i32, i32, vp, vp]),
    "mgdt_conv2d_path": (C.c_int, [C.POINTER(ConvArgs)]),
    "mgdt_conv2d": (C.c_int, [C.POINTER(ConvArgs), vp]),
    "mgdt_conv_umma_packed_bytes": (sz, [i32, i32, i32, i32]),
    "mgdt_conv_umma_pack": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, vp, vp]),
    "mgdt_conv_umma_pack_scaled": (C.c_int, [vp, i32, i32, i32, i32, vp, i32, vp, vp]),
    "mgdt_conv_umma_pack_scaled_groups": (C.c_int, [vp, i32, i32, i32, i32, vp, i32, i32, i32, vp, vp]),
    "mgdt_stem_conv": (C.c_int, [vp, i32, vp, i32, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    "mgdt_stem_u8_supported": (C.c_int, [i32, i32, i32, i32, i32]),
    "mgdt_stem_u8_packed_bytes": (sz, [i32]),
    "mgdt_stem_u8_pack": (C.c_int, [vp, i32, i32, vp, vp]),
    "mgdt_stem_u8": (C.c_int, [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "mgdt_tood_cls_supported": (C.c_int, [i32, i32, i32, i32, i32]),
    "mgdt_tood_cls": (C.c_int, [vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    "mgdt_mspa_front_supported": (C.c_int, [i32, i32]),
    "mgdt_mspa_front_packed_bytes": (sz, [i32, i32]),
    "mgdt_mspa_front_pack": (C.c_int, [vp, i32, i32, vp, vp]),
    "mgdt_mspa_front": (C.c_int, [vp, i32, vp, vp, i32, i32, i32, vp, i32, vp, i32, i32, i32, i32, i32, vp]),
    "mgdt_dwconv7_ln": (C.c_int, [vp, i32, vp, vp, vp, vp, f32, vp, i32, i32, i32, i32, i32, i32, vp]),
    "mgdt_dcn3x3": (C.c_int, [vp, i32, vp, i32, vp, i32, i32, vp, vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, vp, i32, i32, i32, vp]),
    "mgdt_dcn3x3_path": (C.c_int, [vp, i32, vp, i32, i32, i32, i32, i32, i32]),
    "mgdt_chan_stats_ws_bytes": (sz, [i32, i32, i32, i32, i32]),
    "mgdt_chan_stats": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, sz, vp, i32, vp]),
    "mgdt_chan_stats_fin": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, sz, vp, C.POINTER(StatsFin), i32, vp]),
    "mgdt_stats_finish": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, C.POINTER(StatsFin), vp]),
    "mgdt_mspa_gate": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, i32, vp, vp]),
    "mgdt_grn_scale": (C.c_int, [vp, vp, i32, i32, vp, vp]),
    "mgdt_gn_affine": (C.c_int, [vp, vp, i32, i32, i32, i32, f32, vp, vp, vp, vp, vp]),
    "mgdt_td_attn": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]),
    "mgdt_affine_act": (C.c_int, [vp, i32, vp, vp, vp, i32, i32, vp, i32, i32, i32, i32, i32, i32, vp]),
    "mgdt_resample": (C.c_int, [vp, i32, i32, i32, vp, i32, i32, i32, i32, i32, i32, i32, vp]),
    "mgdt_sppf_pool": (C.c_int, [vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]),
    "mgdt_inject": (C.c_int, [vp, i32, vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    "mgdt_inject2": (C.c_int, [vp, i32, vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    "mgdt_preprocess": (C.c_int, [vp, i32, vp, i32, i32, i32, i32, i32, i32, vp]),
    "mgdt_decode": (C.c_int, [C.POINTER(DecodeLevel), i32, i32, i32, i32, i32, vp, i32, vp]),
    "mgdt_nms_ws_bytes": (sz, [i32, i32, i32, i32, i32]),
    "mgdt_nms": (C.c_int, [vp, i32, i32, i32, f32, f32, i32, i32, i32, i32, f32, vp, i32, vp, vp, vp, sz, vp]),
    "mgdt_v8_loss_ws_bytes": (sz, [i32, i32, i32]),
    "mgdt_v8_loss": (C.c_int, [vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, f32, i32, f32, f32, f32, vp, vp, vp, vp, vp, vp, vp, sz, vp]),
    "mgdt_ema_update": (C.c_int, [vp, vp, sz, f32, vp]),
    "mgdt_sumsq": (C.c_int, [vp, sz, vp, vp]),
    "mgdt_sgd_step": (C.c_int, [vp, vp, vp, vp, sz, C.POINTER(C.c_float), C.POINTER(C.c_float), f32, i32, i32, vp, f32, f32, vp]),
}

_LIB = None


def lib():
    """Load (once) and return the CDLL; raises if the library was not built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"mgdt_yolo_b200: {LIB_PATH} is missing -- build it with `python -m mgdt_yolo_b200.build` "
                "(nvcc, sm_100a). There is no CPU or PyTorch fallback for this path.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the symbol is missing
            fn.restype, fn.argtypes = res, args
        if L.mgdt_abi_version() != ABI_VERSION:
            raise RuntimeError("mgdt_yolo_b200: ABI version mismatch, rebuild the library")
        # the library reads no environment variables: forward the documented MGDT_* switches once, here
        for env, opt in (("MGDT_PDL", "pdl"), ("MGDT_CONV_TMA_LOAD", "conv_tma_load"), ("MGDT_CONV_TMA_STORE", "conv_tma_store"), ("MGDT_CONV_TMA_STATS", "conv_tma_stats"), ("MGDT_CONV_KSPLIT", "conv_ksplit"),
                         ("MGDT_CONV_PAIR", "conv_pair"), ("MGDT_CONV_SPLIT", "conv_split"), ("MGDT_CONV_MB", "conv_mb"), ("MGDT_CONV_TMA3X3", "conv_tma3x3"), ("MGDT_CONV_TMA3X3_S2", "conv_tma3x3_s2"), ("MGDT_DW_PAIRS", "dw_pairs"), ("MGDT_CONV3X3_WARP", "conv3x3_warp"), ("MGDT_CW_SPC", "conv3x3_warp_spc")):
            v = os.environ.get(env)
            if v is not None and v.lstrip("-").isdigit():
                L.mgdt_set_option(opt.encode(), int(v))
        _LIB = L
    return _LIB


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().mgdt_last_error().decode(errors="replace")
        raise RuntimeError(f"mgdt_b200 {what} failed ({rc}): {msg}")
