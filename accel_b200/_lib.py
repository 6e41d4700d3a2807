"""ctypes binding of include/accel_b200.h.  There is no fallback: if the library is missing or the
process has no CUDA device, the compute entry points raise."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaccel_b200.so")

# every symbol include/accel_b200.h declares
SYMBOLS = (
    "accel_create", "accel_destroy", "accel_last_error", "accel_param_count", "accel_param_info",
    "accel_set_param", "accel_finalize", "accel_key_forward", "accel_cur_forward", "accel_key_forward_lin",
    "accel_cur_forward_lin", "accel_flownet", "accel_rbranch_forward", "accel_plan_interval", "accel_interval_forward",
    "accel_graph_cache_stats", "accel_debug_fetch",
    "accel_warp", "accel_warp_split", "accel_fuse_argmax", "accel_preprocess", "accel_resize_size", "accel_resize_bgr", "accel_confusion", "accel_conv_layer", "accel_head", "accel_last_launch_count",
    "accel_set_profiling", "accel_stage_times", "accel_op_times",
)


class AccelConfig(C.Structure):
    _fields_ = [("version", C.c_int), ("height", C.c_int), ("width", C.c_int), ("num_classes", C.c_int),
                ("device", C.c_int), ("flags", C.c_int)]


FLAG_NO_TENSOR_CORES = 1
FLAG_NO_GRAPH = 2
VERSION_CODE = {"dff": 0, "18": 18, "34": 34, "50": 50, "101": 101}

_lib = None


def load():
    """Loads libaccel_b200.so (built by `python accel_b200/build.py` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("ACCEL_B200_LIB", LIB_PATH)          # A/B aid: another build of the same ABI
    if not os.path.exists(path):
        raise RuntimeError("accel_b200: %s is missing -- build it with `python accel_b200/build.py` "
                           "(there is no CPU or PyTorch fallback)" % path)
    lib = C.CDLL(path)
    vp, cp, ip, fp, u8p = C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_float), C.c_void_p
    lib.accel_create.argtypes = [C.POINTER(AccelConfig), C.POINTER(vp)]
    lib.accel_destroy.argtypes = [vp]
    lib.accel_destroy.restype = None
    lib.accel_last_error.argtypes = [vp]
    lib.accel_last_error.restype = cp
    lib.accel_param_count.argtypes = [vp]
    lib.accel_param_info.argtypes = [vp, ip, C.POINTER(cp), C.POINTER(C.c_int64), C.POINTER(C.c_int)]
    lib.accel_set_param.argtypes = [vp, cp, vp, C.POINTER(C.c_int64), ip]
    lib.accel_finalize.argtypes = [vp]
    lib.accel_key_forward.argtypes = [vp, vp, vp, vp, u8p, vp]
    lib.accel_cur_forward.argtypes = [vp, vp, vp, vp, vp, vp, u8p, vp]
    lib.accel_key_forward_lin.argtypes = [vp, vp, vp, vp, vp, u8p, vp]
    lib.accel_cur_forward_lin.argtypes = [vp, vp, vp, vp, vp, vp, u8p, vp]
    lib.accel_flownet.argtypes = [vp, vp, vp, vp, vp]
    lib.accel_rbranch_forward.argtypes = [vp, vp, vp, u8p, vp]
    lib.accel_plan_interval.argtypes = [vp, ip]
    lib.accel_interval_forward.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), vp]
    lib.accel_debug_fetch.argtypes = [vp, cp, cp, vp, C.POINTER(C.c_int64), vp]
    lib.accel_graph_cache_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.accel_warp.argtypes = [vp, vp, vp, ip, ip, ip, vp]
    lib.accel_warp_split.argtypes = [vp, vp, vp, vp, vp, ip, ip, ip, vp]
    lib.accel_fuse_argmax.argtypes = [vp, vp, vp, vp, ip, ip, ip, u8p, vp, vp]
    lib.accel_preprocess.argtypes = [u8p, ip, ip, C.POINTER(C.c_double), vp, vp]
    lib.accel_resize_size.argtypes = [ip, ip, C.c_double, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.accel_resize_bgr.argtypes = [u8p, ip, ip, C.c_double, C.c_double, u8p, vp]
    lib.accel_confusion.argtypes = [u8p, u8p, C.c_size_t, ip, vp, vp]
    lib.accel_conv_layer.argtypes = [ip, vp, ip, ip, ip, vp, ip, ip, ip, ip, ip, ip, vp, vp, vp, ip, vp, ip, vp, ip,
                                     C.c_char_p, ip]
    lib.accel_head.argtypes = [vp, ip, ip, ip, vp, vp, ip, vp, vp, ip, vp, ip, C.c_char_p, ip]
    lib.accel_last_launch_count.argtypes = [vp]
    lib.accel_set_profiling.argtypes = [vp, ip]
    lib.accel_stage_times.argtypes = [vp, C.POINTER(cp), fp, ip]
    lib.accel_op_times.argtypes = [vp, C.POINTER(cp), fp, C.POINTER(C.c_double), ip]
    _lib = lib
    return lib
