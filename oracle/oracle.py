"""ctypes front-end to the oracle libraries (TEST INFRASTRUCTURE ONLY).

  liboracle.so    -- our plain-C restatement, oracle/stereo_oracle.c (always available after
                     `make -C oracle`)
  _ref/libref.so  -- the reference's own BlockMatching.cpp + ctmf.c compiled UNMODIFIED from
                     /root/reference (built in the dev container, travels as a prebuilt .so)
  _ref/libstref.so -- likewise STMatching/StereoHelper.cpp (float WTA, right-view cost volume)
  _ref/libutilref.so -- likewise BlockMatching/Utility.cpp (CPU_Remap, cvtColor_cpu)

Every wrapper cites the reference file:line its C counterpart follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(verbose: bool = False) -> None:
    """make -C oracle (liboracle.so always; _ref/libref.so when /root/reference exists)."""
    out = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


_lib = None
_ref = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_ad_slice.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p]
        L.orc_ad_volume.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, _u8p]
        L.orc_sad_slice.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p]
        L.orc_sad_wta_direct.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p]
        L.orc_sad_wta.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_void_p]
        L.orc_all_sad.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p]
        L.orc_gf_cost_slices.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                         C.c_int, C.c_int, _f64p]
        L.orc_gf_wta.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                 _u8p, C.c_void_p]
        L.orc_lr_check.argtypes = [_u8p, _u8p, C.c_int, C.c_int, _u8p, _u8p]
        L.orc_median.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int]
        _f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
        L.orc_remap.argtypes = [_u8p, _f32p, _f32p, C.c_int, C.c_int, _u8p]
        L.orc_remap.restype = None
        L.orc_cvtcolor.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p]
        L.orc_cvtcolor.restype = None
        L.orc_fnv1a64.argtypes = [_u8p, C.c_size_t]
        L.orc_fnv1a64.restype = C.c_uint64
        for f in ("orc_ad_slice", "orc_ad_volume", "orc_sad_slice", "orc_sad_wta_direct", "orc_sad_wta",
                  "orc_all_sad", "orc_gf_cost_slices", "orc_gf_wta", "orc_lr_check", "orc_median"):
            getattr(L, f).restype = None
        _lib = L
    return _lib


def have_ref() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libref.so"))


def ref() -> C.CDLL:
    """The reference's own CPU code, compiled unmodified (oracle/Makefile target _ref/libref.so)."""
    global _ref
    if _ref is None:
        path = os.path.join(_HERE, "_ref", "libref.so")
        if not os.path.exists(path):
            build()
        R = C.CDLL(path)
        for f in ("ref_PreCal", "ref_getDisp", "ref_getAllSAD"):
            getattr(R, f).argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p]
            getattr(R, f).restype = None
        R.ref_compareDisp.argtypes = [_u8p, _u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int]
        R.ref_compareDisp.restype = None
        R.ref_median.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int]
        R.ref_median.restype = None
        _ref = R
    return _ref


_stref = None


def have_stref() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libstref.so"))


def stref() -> C.CDLL:
    """The reference's own STMatching/StereoHelper.cpp, compiled unmodified (oracle/Makefile target
    _ref/libstref.so): ref_wta_float (StereoHelper.cpp:131-154), ref_right_from_left (:156-180)."""
    global _stref
    if _stref is None:
        S = C.CDLL(os.path.join(_HERE, "_ref", "libstref.so"))
        f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
        S.ref_wta_float.argtypes = [f32, C.c_int, C.c_int, C.c_int, _u8p]
        S.ref_wta_float.restype = None
        S.ref_right_from_left.argtypes = [f32, C.c_int, C.c_int, C.c_int, f32]
        S.ref_right_from_left.restype = None
        _stref = S
    return _stref


_utilref = None


def have_utilref() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libutilref.so"))


def utilref() -> C.CDLL:
    """The reference's own BlockMatching/Utility.cpp, compiled unmodified (oracle/Makefile target
    _ref/libutilref.so): ref_cpu_remap (Utility.cpp:236-264), ref_cvtcolor_cpu (:289-298)."""
    global _utilref
    if _utilref is None:
        U = C.CDLL(os.path.join(_HERE, "_ref", "libutilref.so"))
        f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
        U.ref_cpu_remap.argtypes = [_u8p, C.c_int, C.c_int, f32, f32, _u8p]
        U.ref_cpu_remap.restype = None
        U.ref_cvtcolor_cpu.argtypes = [_u8p, _u8p, C.c_int, C.c_int]
        U.ref_cvtcolor_cpu.restype = None
        _utilref = U
    return _utilref


_segref = None


def have_segref() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libsegref.so"))


def segref() -> C.CDLL:
    """The reference's own segment-tree stereo -- STMatching/SegmentTree.cpp, segment-graph.h, disjoint-set.h,
    StereoHelper.cpp, StereoDisparity.cpp, Toolkit.cpp, ctmf.c compiled unmodified (oracle/Makefile target
    _ref/libsegref.so, driver oracle/seg_driver.cpp)."""
    global _segref
    if _segref is None:
        S = C.CDLL(os.path.join(_HERE, "_ref", "libsegref.so"))
        f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
        S.ref_st_matching_cost.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, f32]
        S.ref_st_matching_cost.restype = None
        S.ref_st_filter.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]
        S.ref_st_filter.restype = None
        for f in ("ref_st_routine", "ref_st_iteration"):
            getattr(S, f).argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _u8p]
            getattr(S, f).restype = None
        _segref = S
    return _segref


def _bgr(a) -> np.ndarray:
    a = np.ascontiguousarray(a, np.uint8)
    assert a.ndim == 3 and a.shape[2] == 3
    return a


def ref_st_matching_cost(left_bgr, right_bgr, D: int) -> np.ndarray:
    """GetMatchingCost of the compiled reference (StereoHelper.cpp:75-129): float32 [H][W][D]."""
    L, R = _bgr(left_bgr), _bgr(right_bgr)
    out = np.empty(L.shape[:2] + (D,), np.float32)
    segref().ref_st_matching_cost(L.reshape(-1), R.reshape(-1), L.shape[1], L.shape[0], D, out)
    return out


def ref_st_filter(image_bgr, cost=None, sigma: float = 0.1, tau: float = 1200.0):
    """CColorWeight + BuildSegmentTree (+ Filter on a copy of cost, float32 [H][W][D]) of the compiled reference
    (SegmentTree.cpp:38-195): (aggregated volume | None, order, father_id, father_dist)."""
    img = _bgr(image_bgr)
    h, w = img.shape[:2]
    n = h * w
    order, father, fdist = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.uint8)
    vol = None if cost is None else np.ascontiguousarray(cost, np.float32).copy()
    D = 1 if vol is None else vol.shape[2]
    segref().ref_st_filter(img.reshape(-1), w, h, D, sigma, tau, None if vol is None else vol.ctypes.data,
                           order.ctypes.data, father.ctypes.data, fdist.ctypes.data)
    return vol, order, father, fdist


def ref_st_routine(left_bgr, right_bgr, D: int, scale: int = 1, sigma: float = 0.1, refined: bool = False) -> np.ndarray:
    """stereo_disparity_normal (StereoDisparity.cpp:58-90) or, refined=True, stereo_disparity_iteration (:92-160) of the
    compiled reference: u8 [H][W]."""
    L, R = _bgr(left_bgr), _bgr(right_bgr)
    out = np.empty(L.shape[:2], np.uint8)
    fn = segref().ref_st_iteration if refined else segref().ref_st_routine
    fn(L.reshape(-1), R.reshape(-1), L.shape[1], L.shape[0], D, scale, sigma, out)
    return out


def st_edge_weights(image_bgr):
    """CColorWeight (SegmentTree.cpp:183-195) restated: 3x3 median per channel (ctmf r=1 == replicate-border median), then
    max over the channels of |a - b| between 4-neighbours: (wr, wu) u8 [H][W], 255 where the edge does not exist."""
    img = _bgr(image_bgr)
    med = np.stack([median(np.ascontiguousarray(img[:, :, c]), 1) for c in range(3)], -1).astype(np.int32)
    wr = np.full(img.shape[:2], 255, np.uint8)
    wu = np.full(img.shape[:2], 255, np.uint8)
    wr[:, :-1] = np.abs(med[:, :-1] - med[:, 1:]).max(-1)
    wu[1:, :] = np.abs(med[1:, :] - med[:-1, :]).max(-1)
    return wr, wu


def ref_cpu_remap(src, mapx, mapy) -> np.ndarray:
    """CPU_Remap of the compiled reference (Utility.cpp:236-246)."""
    src = _u8(src)
    out = np.empty_like(src)
    utilref().ref_cpu_remap(src, src.shape[0], src.shape[1], np.ascontiguousarray(mapx, np.float32),
                            np.ascontiguousarray(mapy, np.float32), out)
    return out


def ref_cvtcolor_cpu(src3) -> np.ndarray:
    """cvtColor_cpu of the compiled reference (Utility.cpp:289-298)."""
    src3 = np.ascontiguousarray(src3, np.uint8)
    out = np.empty(src3.shape[:2], np.uint8)
    utilref().ref_cvtcolor_cpu(src3.reshape(-1), out, src3.shape[0], src3.shape[1])
    return out


def ref_wta_float(vol) -> np.ndarray:
    """GetDisparity_WTA of the compiled reference on a float32 volume [H][W][D]."""
    vol = np.ascontiguousarray(vol, np.float32)
    h, w, d = vol.shape
    out = np.empty((h, w), np.uint8)
    stref().ref_wta_float(vol, w, h, d, out)
    return out


def ref_right_from_left(vol) -> np.ndarray:
    """GetRightMatchingCostFromLeft of the compiled reference on a float32 volume [H][W][D]."""
    vol = np.ascontiguousarray(vol, np.float32)
    h, w, d = vol.shape
    out = np.empty_like(vol)
    stref().ref_right_from_left(vol, w, h, d, out)
    return out


@contextlib.contextmanager
def quiet_stdout():
    """The reference prints phase timings with cout (BlockMatching.cpp:136,153,188); silence fd 1."""
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(devnull, 1)
        yield
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)


def _u8(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2
    return a


# ----------------------------------------------------------------------------- restatement
def ad_slice(L, R, d: int, view: int = 0) -> np.ndarray:
    """A.1: BlockMatching.cpp:101-108 (view 0); StereoHelper.cpp:156-180 (view 1)."""
    L, R = _u8(L), _u8(R)
    out = np.empty_like(L)
    lib().orc_ad_slice(L, R, L.shape[0], L.shape[1], d, view, out)
    return out


def ad_volume(L, R, D: int) -> np.ndarray:
    """PreCal, BlockMatching.cpp:89-109 -> u8 [D][H][W]."""
    L, R = _u8(L), _u8(R)
    out = np.empty((D,) + L.shape, np.uint8)
    lib().orc_ad_volume(L, R, L.shape[0], L.shape[1], D, out)
    return out


def sad_slice(L, R, r: int, d: int, view: int = 0) -> np.ndarray:
    """Un-truncated clipped-window SAD (BlockMatching.cpp:167-177) -> int32 [H][W]."""
    L, R = _u8(L), _u8(R)
    out = np.empty(L.shape, np.int32)
    lib().orc_sad_slice(L, R, L.shape[0], L.shape[1], r, d, view, out)
    return out


def sad_wta(L, R, r: int, D: int, direct: bool = False, return_cost: bool = False):
    """getDisp, BlockMatching.cpp:111-189 (A.2)."""
    L, R = _u8(L), _u8(R)
    out = np.empty_like(L)
    if direct:
        lib().orc_sad_wta_direct(L, R, L.shape[0], L.shape[1], r, D, out)
        return out
    cost = np.empty(L.shape, np.int32) if return_cost else None
    lib().orc_sad_wta(L, R, L.shape[0], L.shape[1], r, D, out,
                      cost.ctypes.data_as(C.c_void_p) if return_cost else None)
    return (out, cost) if return_cost else out


def all_sad(L, R, r: int, D: int) -> np.ndarray:
    """getAllSAD, BlockMatching.cpp:191-261 -> u8 [H*W][D] (truncated, 255 where x+d > W)."""
    L, R = _u8(L), _u8(R)
    out = np.empty((L.size, D), np.uint8)
    lib().orc_all_sad(L, R, L.shape[0], L.shape[1], r, D, out)
    return out


GF_EPS_DEFAULT = 1e-4 * 255.0 * 255.0  # 6.5025


def gf_cost_slices(L, R, r: int, d0: int, nd: int, eps: float = GF_EPS_DEFAULT, view: int = 0) -> np.ndarray:
    """GF-v1 aggregated costs q_d, float64 [nd][H][W] (A.3; parity unpinned)."""
    L, R = _u8(L), _u8(R)
    out = np.empty((nd,) + L.shape, np.float64)
    lib().orc_gf_cost_slices(L, R, L.shape[0], L.shape[1], r, eps, view, d0, nd, out)
    return out


def gf_wta(L, R, r: int, D: int, eps: float = GF_EPS_DEFAULT, view: int = 0, return_cost: bool = False):
    """GF-v1 + float WTA (A.3 + A.4, StereoHelper.cpp:137-150)."""
    L, R = _u8(L), _u8(R)
    out = np.empty_like(L)
    cost = np.empty(L.shape, np.float64) if return_cost else None
    lib().orc_gf_wta(L, R, L.shape[0], L.shape[1], r, D, eps, view, out,
                     cost.ctypes.data_as(C.c_void_p) if return_cost else None)
    return (out, cost) if return_cost else out


def lr_check(DL, DR):
    """StereoDisparity.cpp:136-147 -> (occ, mask)."""
    DL, DR = _u8(DL), _u8(DR)
    occ = np.empty_like(DL)
    mask = np.empty_like(DL)
    lib().orc_lr_check(DL, DR, DL.shape[0], DL.shape[1], occ, mask)
    return occ, mask


def median(img, r: int) -> np.ndarray:
    """ctmf semantics (ctmf.c:378-433): (2r+1)^2 median, replicate border."""
    img = _u8(img)
    out = np.empty_like(img)
    lib().orc_median(img, out, img.shape[0], img.shape[1], r)
    return out


def remap(src, mapx, mapy) -> np.ndarray:
    """CPU_Remap / kernalRemap (Utility.cpp:236-264, Device.cu:127-167): bilinear, OOB -> 0, round-nearest-even."""
    src = _u8(src)
    mx = np.ascontiguousarray(mapx, np.float32); my = np.ascontiguousarray(mapy, np.float32)
    out = np.empty_like(src)
    lib().orc_remap(src, mx, my, src.shape[0], src.shape[1], out)
    return out


def cvtcolor(src3, truncate: bool = False) -> np.ndarray:
    """kernalCvtColor (round, Device.cu:136-143) / cvtColor_cpu (truncate, Utility.cpp:289-298)."""
    a = np.ascontiguousarray(src3, np.uint8)
    assert a.ndim == 3 and a.shape[2] == 3
    out = np.empty(a.shape[:2], np.uint8)
    lib().orc_cvtcolor(a.reshape(-1, 3).reshape(a.shape[0], -1), a.shape[0], a.shape[1], int(truncate), out)
    return out


def fnv1a64(a) -> str:
    a = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    return "%016x" % lib().orc_fnv1a64(a, a.size)


def stereo_pipeline(L, R, *, mode: str, r: int, D: int, eps: float = GF_EPS_DEFAULT,
                    lr: bool = False, median_r: int = 0):
    """Whole path as the product runs it (DESIGN.md "Pipeline order"):
    WTA (left) [-> WTA (right) when lr] -> median on both -> LR check -> occluded pixels zeroed.
    Returns (disp, mask) ; mask is None without lr.  Order follows StereoDisparity.cpp:115-147.
    """
    if mode == "sad":
        dl = sad_wta(L, R, r, D)
    else:
        dl = gf_wta(L, R, r, D, eps, 0)
    if median_r > 0:
        dl = median(dl, median_r)
    if not lr:
        return dl, None
    if mode == "sad":
        raise ValueError("the reference defines no right-view SAD; lr needs mode='gf'")
    dr = gf_wta(L, R, r, D, eps, 1)
    if median_r > 0:
        dr = median(dr, median_r)
    occ, mask = lr_check(dl, dr)
    out = dl.copy()
    out[occ != 0] = 0
    return out, mask


# ----------------------------------------------------------------------------- real reference
def ref_getDisp(L, R, r: int, D: int) -> np.ndarray:
    L, R = _u8(L), _u8(R)
    out = np.empty_like(L)
    with quiet_stdout():
        ref().ref_getDisp(L, R, L.shape[0], L.shape[1], r, D, out)
    return out


def ref_PreCal(L, R, D: int) -> np.ndarray:
    L, R = _u8(L), _u8(R)
    out = np.zeros((D,) + L.shape, np.uint8)
    ref().ref_PreCal(L, R, L.shape[0], L.shape[1], 0, D, out)
    return out


def ref_getAllSAD(L, R, r: int, D: int) -> np.ndarray:
    L, R = _u8(L), _u8(R)
    out = np.full((L.size, D), 255, np.uint8)
    with quiet_stdout():
        ref().ref_getAllSAD(L, R, L.shape[0], L.shape[1], r, D, out)
    return out


def ref_median(img, r: int) -> np.ndarray:
    img = _u8(img)
    out = np.empty_like(img)
    ref().ref_median(img, out, img.shape[0], img.shape[1], r)
    return out
