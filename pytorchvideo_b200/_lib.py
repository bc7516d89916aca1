"""ctypes binding of libpvb200.so (the C ABI declared in include/pv_b200.h).

There is deliberately NO fallback: if the library is missing or a call fails, a RuntimeError is
raised.  PyTorch is used only for device memory and streams; nothing here calls ATen compute.
"""
import ctypes as C
import os

from . import _build

PV_F16, PV_F32, PV_U8 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_SWISH, ACT_GELU, ACT_SIGMOID = 0, 1, 2, 3, 4
ALGO_AUTO, ALGO_DIRECT, ALGO_TCGEN05 = 0, 1, 2
POOL_MAX, POOL_AVG = 0, 1

c_ll = C.c_longlong
c_vp = C.c_void_p


class ClipTransformDesc(C.Structure):
    _fields_ = [("C", C.c_int), ("n_t", C.c_int), ("out_h", C.c_int), ("out_w", C.c_int),
                ("sc", c_ll), ("st", c_ll), ("sh", c_ll), ("sw", c_ll),
                ("mean", C.c_float * 4), ("stdv", C.c_float * 4),
                ("src_dtype", C.c_int), ("dst_dtype", C.c_int), ("div255", C.c_int)]


class ClipBatchDesc(C.Structure):
    _fields_ = [("C", C.c_int), ("n_clips", C.c_int), ("n_t", C.c_int), ("n_slow", C.c_int),
                ("in_h", C.c_int), ("in_w", C.c_int), ("new_h", C.c_int), ("new_w", C.c_int),
                ("top", C.c_int), ("left", C.c_int), ("out_h", C.c_int), ("out_w", C.c_int),
                ("hflip", C.c_int),
                ("sc", c_ll), ("st", c_ll), ("sh", c_ll), ("sw", c_ll), ("s_clip", c_ll),
                ("d_clip", c_ll), ("d_slow_clip", c_ll),
                ("mean", C.c_float * 4), ("stdv", C.c_float * 4),
                ("div255", C.c_int), ("normalize", C.c_int), ("src_dtype", C.c_int), ("dst_dtype", C.c_int)]


class BottleneckDesc(C.Structure):
    _fields_ = [("N", C.c_int), ("T", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("Cin", C.c_int), ("Cmid", C.c_int), ("Cout", C.c_int), ("kt", C.c_int), ("sb", C.c_int),
                ("has_shortcut", C.c_int), ("act", C.c_int), ("x_row_stride", c_ll), ("y_row_stride", c_ll)]


class Conv3dDesc(C.Structure):
    _fields_ = [("dtype", C.c_int),
                ("N", C.c_int), ("Ti", C.c_int), ("Hi", C.c_int), ("Wi", C.c_int), ("Ci", C.c_int),
                ("To", C.c_int), ("Ho", C.c_int), ("Wo", C.c_int), ("Co", C.c_int),
                ("kt", C.c_int), ("kh", C.c_int), ("kw", C.c_int),
                ("st", C.c_int), ("sh", C.c_int), ("sw", C.c_int),
                ("pt", C.c_int), ("ph", C.c_int), ("pw", C.c_int),
                ("dt", C.c_int), ("dh", C.c_int), ("dw", C.c_int),
                ("groups", C.c_int), ("act", C.c_int), ("has_residual", C.c_int),
                ("x_row_stride", c_ll), ("y_row_stride", c_ll), ("res_row_stride", c_ll),
                ("ci_pad64", C.c_int), ("x_w_pad", C.c_int), ("x_w_phys", C.c_int),
                ("x_batch_stride", c_ll), ("y_batch_stride", c_ll)]


class Pool3dDesc(C.Structure):
    _fields_ = [("dtype", C.c_int), ("mode", C.c_int),
                ("N", C.c_int), ("Ti", C.c_int), ("Hi", C.c_int), ("Wi", C.c_int), ("C", C.c_int),
                ("To", C.c_int), ("Ho", C.c_int), ("Wo", C.c_int),
                ("kt", C.c_int), ("kh", C.c_int), ("kw", C.c_int),
                ("st", C.c_int), ("sh", C.c_int), ("sw", C.c_int),
                ("pt", C.c_int), ("ph", C.c_int), ("pw", C.c_int),
                ("x_row_stride", c_ll), ("y_row_stride", c_ll),
                ("x_batch_stride", c_ll), ("y_batch_stride", c_ll)]


class AttentionDesc(C.Structure):
    _fields_ = [("dtype", C.c_int), ("B", C.c_int), ("H", C.c_int), ("Nq", C.c_int),
                ("Nk", C.c_int), ("D", C.c_int),
                ("q_row_stride", c_ll), ("k_row_stride", c_ll), ("v_row_stride", c_ll),
                ("o_row_stride", c_ll),
                ("q_batch_stride", c_ll), ("k_batch_stride", c_ll), ("v_batch_stride", c_ll),
                ("o_batch_stride", c_ll),
                ("scale", C.c_float), ("add_q_residual", C.c_int)]


# name -> (restype, argtypes); mirrors include/pv_b200.h one to one (tests check this list
# against the header and against the symbols the .so exports).
SIGNATURES = {
    "pv_abi_version": (C.c_int, []),
    "pv_last_error": (C.c_char_p, []),
    "pv_device_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pv_launch_count": (c_ll, []),
    "pv_clip_transform_fwd": (C.c_int, [C.POINTER(ClipTransformDesc), c_vp, c_vp, c_vp, c_vp, c_vp,
                                        c_vp, c_vp, c_vp, c_vp, c_vp]),
    "pv_clip_transform_batch": (C.c_int, [C.POINTER(ClipBatchDesc), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "pv_view_reduce": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, c_vp]),
    "pv_ncdhw_to_ndhwc": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.c_int, c_ll, c_vp]),
    "pv_ncdhw_to_ndhwc_padw": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_vp]),
    "pv_zero_f32": (C.c_int, [c_vp, c_ll, c_vp]),
    "pv_ndhwc_to_ncdhw": (C.c_int, [c_vp, C.c_int, c_ll, c_vp, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_int, c_vp]),
    "pv_conv3d_fwd": (C.c_int, [C.POINTER(Conv3dDesc), C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                                c_vp, c_vp]),
    "pv_dwconv3d_fwd": (C.c_int, [C.POINTER(Conv3dDesc), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "pv_temporal_tap_sum": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, c_ll, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int, c_vp, c_vp, C.c_int, c_ll, c_ll, c_vp]),
    "pv_conv3d_tcgen05_supported": (C.c_int, [C.POINTER(Conv3dDesc)]),
    "pv_conv3d_stem_rows_supported": (C.c_int, [C.POINTER(Conv3dDesc)]),
    "pv_conv3d_stem_rows_fwd": (C.c_int, [C.POINTER(Conv3dDesc), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "pv_bottleneck_fused_supported": (C.c_int, [C.POINTER(BottleneckDesc)]),
    "pv_bottleneck_fused_fwd": (C.c_int, [C.POINTER(BottleneckDesc)] + [c_vp] * 15),
    "pv_pool3d_fwd": (C.c_int, [C.POINTER(Pool3dDesc), c_vp, c_vp, c_vp]),
    "pv_channel_sum": (C.c_int, [c_vp, C.c_int, c_ll, C.c_int, c_ll, C.c_int, c_vp, c_vp]),
    "pv_se_gate": (C.c_int, [c_vp, c_ll, C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp,
                             C.c_int, c_vp, c_vp]),
    "pv_scale_act": (C.c_int, [c_vp, c_vp, C.c_int, c_ll, c_ll, C.c_int, c_ll, C.c_int, c_vp,
                               C.c_int, c_vp]),
    "pv_head_reduce": (C.c_int, [c_vp, C.c_int, c_ll, C.c_int, c_ll, C.c_int, C.c_int, c_vp,
                                 c_vp]),
    "pv_roi_align_fwd": (C.c_int, [c_vp, C.c_int, c_ll, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, C.c_int, C.c_int,
                                   C.c_int, C.c_float, C.c_int, c_vp, c_ll, c_vp]),
    "pv_layernorm": (C.c_int, [c_vp, c_vp, C.c_int, c_ll, C.c_int, C.c_int, c_ll, c_ll, c_vp, c_vp,
                               C.c_float, c_vp]),
    "pv_layernorm_sets": (C.c_int, [c_vp, c_vp, C.c_int, c_ll, C.c_int, C.c_int, c_ll, c_ll, c_vp, c_vp, C.c_int, c_vp,
                                    c_ll, c_ll, C.c_float, c_vp]),
    "pv_copy_rows": (C.c_int, [c_vp, c_vp, C.c_int, c_ll, C.c_int, c_ll, c_ll, c_vp]),
    "pv_add_pos_cls": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int, c_ll, C.c_int, c_ll, c_vp, C.c_int, c_vp]),
    "pv_add_pos_cls_to": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int, C.c_int, c_ll, C.c_int, c_ll, c_vp, C.c_int, c_vp]),
    "pv_add_layernorm": (C.c_int, [c_vp, C.c_int, c_ll, c_vp, c_ll, c_vp, c_ll, c_vp, c_ll, c_ll, C.c_int, c_vp, c_vp,
                                   C.c_float, c_vp]),
    "pv_attention_fwd": (C.c_int, [C.POINTER(AttentionDesc), c_vp, c_vp, c_vp, c_vp, c_vp]),
}

_lib = None


def lib_path():
    return _build.LIB_PATH


def load():
    """Load (building first if sources are newer) and return the ctypes library handle."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path) or _build.needs_build():
        path = _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.pv_abi_version() != 1:
        raise RuntimeError("libpvb200 ABI version mismatch")
    _lib = lib
    return lib


def last_error():
    return load().pv_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("libpvb200 %s failed (status %d): %s" % (what, rc, last_error()))


def require_device():
    """Raise unless an sm_100 GPU is visible (the product path has no CPU implementation)."""
    sm, cc = C.c_int(0), C.c_int(0)
    rc = load().pv_device_info(C.byref(sm), C.byref(cc))
    check(rc, "pv_device_info")
    return sm.value, cc.value


def launch_count():
    return int(load().pv_launch_count())
