"""ctypes binding of libdepthpro_b200.so (C-ABI declared in include/depthpro_b200.h).

The library is built in-tree by ``ml-depth-pro-video_b200/build.py`` (nvcc, sm_100a).  There is
no fallback: if it is missing or no B200 is visible, the product path raises.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdepthpro_b200.so")            # 16-bit mode = bfloat16 (also serves fp32 mode)
LIB_PATH_FP16 = os.path.join(_HERE, "libdepthpro_b200_fp16.so")  # same sources, 16-bit mode = IEEE half

PREC_FP32, PREC_BF16 = 0, 1
SRC_F32_CHW, SRC_U8_HWC = 0, 1
FOV_NONE, FOV_HEAD_ONLY, FOV_ENCODER = 0, 1, 2
INTERP = {"bilinear": 0, "bicubic": 1}
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2

_vp, _i, _i64 = C.c_void_p, C.c_int, C.c_int64

# name -> (restype, argtypes); mirrors include/depthpro_b200.h one to one
SIGNATURES = {
    "dp_last_error": (C.c_char_p, []),
    "dp_version": (_i, []),
    "dp_act_dtype": (C.c_char_p, []),
    "dp_engine_create": (_i, [_i, _i, _i, C.POINTER(_vp)]),
    "dp_engine_create_ex": (_i, [_i, _i, _i, _i, C.POINTER(_vp)]),
    "dp_engine_destroy": (_i, [_vp]),
    "dp_engine_set_weight": (_i, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i, _i]),
    "dp_engine_missing_weights": (_i, [_vp]),
    "dp_engine_finalize": (_i, [_vp]),
    "dp_preprocess": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "dp_preprocess_ex": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "dp_split": (_i, [_vp, _vp, _i, _vp, _vp]),
    "dp_merge": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "dp_forward": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "dp_infer": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "dp_infer_ex": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "dp_infer_host": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "dp_unproject": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dp_colorize": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "dp_colorize_range": (_i, [_vp, _vp, _i, _i, _vp, _vp, C.c_float, C.c_float, _vp]),
    "dp_ground_normalize": (_i, [_vp, _vp, _i64, C.POINTER(C.c_double), C.c_double, _vp, _vp]),
    "dp_ground_grid_adjust": (_i, [_vp, _vp, _i64, _i, C.c_double, _vp, _vp]),
    "dp_tap": (_i, [_vp, C.c_char_p, _vp, _i64, C.POINTER(_i64), _vp]),
    "dp_gemm_test": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "dp_conv3x3_test": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "dp_attention_test": (_i, [_vp, _i, _vp, _vp, _i, _vp]),
    "dp_kernel_bench": (_i, [_vp, _i, _i, _i, _i, _i, C.POINTER(C.c_float)]),
    "dp_profile_enable": (_i, [_vp, _i]),
    "dp_profile_collect": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_i64)]),
    "dp_debug_counter": (_i64, [_vp, _i, _i]),
    "dp_launch_count": (_i64, [_vp]),
}

_libs = {}


def load(flavour: str = "bf16") -> C.CDLL:
    """Load a flavour of the shared library (once each) and declare every prototype."""
    if flavour not in _libs:
        path = {"bf16": LIB_PATH, "fp16": LIB_PATH_FP16}[flavour]
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} not found: build it with `python ml-depth-pro-video_b200/build.py` "
                "(there is no CPU / PyTorch fallback)")
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.dp_act_dtype().decode() != flavour:
            raise RuntimeError(f"{path} reports 16-bit type {lib.dp_act_dtype().decode()!r}, expected {flavour!r}")
        _libs[flavour] = lib
    return _libs[flavour]


def check(rc: int, lib: C.CDLL = None) -> None:
    """Raise with the failing library's message (each flavour keeps its own thread-local error string)."""
    if rc != 0:
        msgs = [l.dp_last_error().decode(errors="replace") for l in ([lib] if lib is not None else _libs.values())]
        raise RuntimeError("depthpro_b200: " + " | ".join(m for m in msgs if m))


def ptr(t) -> int:
    """Device / host pointer of a tensor, or NULL for None."""
    return None if t is None else t.data_ptr()
