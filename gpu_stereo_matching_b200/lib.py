"""ctypes binding of libgsm.so (C ABI: include/gsm.h).

The CUDA library IS the product: there is no Python / CPU fallback.  Importing this module without
a built libgsm.so raises ImportError with the build command; calling into it without a B200 raises
GsmError from gsm_create.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgsm.so")

GSM_MODE_SAD = 0
GSM_MODE_GF = 1


class GsmParams(C.Structure):
    """struct gsm_params (include/gsm.h)."""
    _fields_ = [
        ("mode", C.c_int),
        ("radius", C.c_int),
        ("num_disp", C.c_int),
        ("eps", C.c_float),
        ("lr_check", C.c_int),
        ("median_radius", C.c_int),
        ("row_bands", C.c_int),
        ("d_begin", C.c_int),
        ("d_end", C.c_int),
        ("rectify", C.c_int),
    ]


class GsmStParams(C.Structure):
    """struct gsm_st_params (include/gsm.h)."""
    _fields_ = [("num_disp", C.c_int), ("sigma", C.c_float), ("tau", C.c_float), ("median_radius", C.c_int),
                ("scale", C.c_int), ("refined", C.c_int)]


class GsmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"gsm error {code}: {msg}")
        self.code = code


# every symbol include/gsm.h declares: (name, restype, argtypes)
_P = C.c_void_p
_PP = C.POINTER(GsmParams)
SYMBOLS = {
    "gsm_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "gsm_destroy": (None, [_P]),
    "gsm_last_error": (C.c_char_p, []),
    "gsm_version": (C.c_char_p, []),
    "gsm_block_matching": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "gsm_stereo_batch": (C.c_int, [_P, _PP, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int]),
    "gsm_stereo_batch_async": (C.c_int, [_P, _PP, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int]),
    "gsm_stereo_device": (C.c_int, [_P, _PP, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "gsm_stereo_batch_v": (C.c_int, [_P, _PP, C.c_int, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P),
                                     C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gsm_stereo_device_v": (C.c_int, [_P, _PP, C.c_int, _P, _P, _P, _P, C.POINTER(C.c_int), C.POINTER(C.c_int), _P]),
    "gsm_postfilter_device": (C.c_int, [_P, _PP, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "gsm_sync": (C.c_int, [_P]),
    "gsm_partial_keys_device": (C.c_int, [_P, _PP, C.c_int, _P, _P, _P, C.c_int, C.c_int, _P]),
    "gsm_partial_keys_device_ex": (C.c_int, [_P, _PP, C.c_int, _P, _P, _P, C.c_int, C.c_int, _P, _P]),
    "gsm_finalize_keys_device": (C.c_int, [_P, _PP, _P, _P, _P, _P, C.c_int, C.c_int, _P]),
    "gsm_reduce_keys_p2p": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.c_int, C.c_int, C.c_longlong, _P]),
    "gsm_reduce_keys_p2p_ex": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.c_int, C.c_int, C.c_longlong, _P, C.c_int]),
    "gsm_ad_volume": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int]),
    "gsm_cost_slices": (C.c_int, [_P, _PP, C.c_int, _P, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int]),
    "gsm_all_sad": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "gsm_median": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int]),
    "gsm_lr_check": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int]),
    "gsm_remap": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int]),
    "gsm_cvtcolor": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int]),
    "gsm_disparity_to_depth": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_float]),
    "gsm_set_rectification": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int]),
    "gsm_segment_tree_stereo": (C.c_int, [_P, C.POINTER(GsmStParams), _P, _P, _P, C.c_int, C.c_int]),
    "gsm_segment_tree_stereo_batch": (C.c_int, [_P, C.POINTER(GsmStParams), _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "gsm_st_matching_cost": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int]),
    "gsm_st_filter": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _P, _P, _P]),
    "gsm_st_build_tree_host": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_float, _P, _P, _P, C.POINTER(C.c_int)]),
    "gsm_launch_count": (C.c_longlong, [_P]),
    "gsm_set_kernel_timing": (C.c_int, [_P, C.c_int]),
    "gsm_last_kernel_ms": (C.c_float, [_P]),
    "gsm_measure_alu_peak": (C.c_int, [_P, C.POINTER(C.c_double)]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C gpu_stereo_matching_b200/csrc` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise GsmError(rc, load().gsm_last_error().decode("utf-8", "replace"))
