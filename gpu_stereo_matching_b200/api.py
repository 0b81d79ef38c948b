"""Host-side mirror of the reference interface for the BlockMatching hot path.

Reference interface (C++, OpenCV types)                      this module
  blockMatching_gpu(Mat&, Mat&, Mat&, int, int)                blockMatching_gpu(left, right, SADWindowSize, searchRange)
      BlockMatching/Device.cuh:50, Device.cu:173-301
  singleFrame()  BlockMatching/Caller.cpp:9-25                 singleFrame(left_gray, right_gray)  (imread/imshow decoupled)
  PreCal / getAllSAD / compareDiff / compareSAD / compareDisp   StereoContext.ad_volume / all_sad / compare_disp
      BlockMatching/BlockMatching.h:8-15
Same argument meaning (SADWindowSize is a radius, searchRange the number of disparities) and the
same result (u8 disparity, rows x cols).  Errors raise GsmError instead of being ignored.

Everything computes in libgsm.so (CUDA, sm_100a) through the C ABI of include/gsm.h.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import lib as _l
from .lib import GSM_MODE_GF, GSM_MODE_SAD, GsmError, GsmParams, GsmStParams  # noqa: F401

GF_EPS_DEFAULT = 6.5025  # 1e-4 * 255^2


def make_params(mode="sad", radius=5, num_disp=64, eps=0.0, lr_check=False, median_radius=0, row_bands=0,
                d_begin=0, d_end=0, rectify=False) -> GsmParams:
    m = {"sad": GSM_MODE_SAD, "gf": GSM_MODE_GF}[mode] if isinstance(mode, str) else int(mode)
    return GsmParams(m, int(radius), int(num_disp), float(eps), int(bool(lr_check)), int(median_radius),
                     int(row_bands), int(d_begin), int(d_end), int(bool(rectify)))


def _u8c(a, name):
    a = np.asarray(a)
    if a.dtype != np.uint8:
        raise TypeError(f"{name}: expected uint8 (CV_8UC1), got {a.dtype}")
    return np.ascontiguousarray(a)  # the reference requires continuous Mats (Device.cu:213-214)


def _same_shape(ndim, **arrays):
    """All arrays C-contiguous uint8 with `ndim` dimensions and ONE shape: the C ABI takes raw pointers plus a single
    rows x cols, so a mismatch would read or write past the end of a host buffer instead of raising."""
    shape = None
    for name, a in arrays.items():
        if a is None:
            continue
        if not isinstance(a, np.ndarray) or a.dtype != np.uint8:
            raise TypeError(f"{name}: expected a uint8 numpy array (CV_8UC1)")
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError(f"{name}: must be C-contiguous")
        if a.ndim != ndim:
            raise ValueError(f"{name}: expected {ndim} dimensions, got shape {a.shape}")
        if shape is None:
            shape = a.shape
        elif a.shape != shape:
            raise ValueError(f"{name}: shape {a.shape} differs from {shape}")
    return shape


def _ptr(a) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(None)


class StereoContext:
    """Owns one gsm_ctx (device buffers, stream) on one GPU.  One host thread at a time."""

    def __init__(self, max_rows: int, max_cols: int, max_disp: int = 256, max_batch: int = 1, device: int = 0):
        self._lib = _l.load()
        self._h = C.c_void_p(None)
        _l.check(self._lib.gsm_create(C.byref(self._h), device, max_rows, max_cols, max_disp, max_batch))
        self.device = device
        self.max_rows, self.max_cols, self.max_disp, self.max_batch = max_rows, max_cols, max_disp, max_batch

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.gsm_destroy(self._h)
            self._h = C.c_void_p(None)

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- reference entry point -------------------------------------------------------------
    def block_matching(self, left, right, radius: int, num_disp: int) -> np.ndarray:
        """== blockMatching_gpu / getDisp (bit-exact), host arrays."""
        L, R = _u8c(left, "left"), _u8c(right, "right")
        if L.ndim != 2 or L.shape != R.shape:
            raise ValueError("left/right must be 2-D and the same size")
        out = np.empty_like(L)
        _l.check(self._lib.gsm_block_matching(self._h, _ptr(L), _ptr(R), _ptr(out), L.shape[0], L.shape[1],
                                              radius, num_disp))
        return out

    # ---- full path -------------------------------------------------------------------------
    def stereo_batch(self, left, right, params: GsmParams, out: Optional[np.ndarray] = None,
                     mask_out: Optional[np.ndarray] = None) -> Tuple[np.ndarray, Optional[np.ndarray]]:
        """n frames [n, rows, cols] (or one frame [rows, cols]) of host u8 -> (disparity, mask|None)."""
        L, R = _u8c(left, "left"), _u8c(right, "right")
        single = L.ndim == 2
        if single:
            L, R = L[None], R[None]
        if L.ndim != 3 or L.shape != R.shape:
            raise ValueError("left/right must be [n, rows, cols] and the same size")
        n, rows, cols = L.shape
        if single and out is not None and out.ndim == 2:
            out = out[None]
        if single and mask_out is not None and mask_out.ndim == 2:
            mask_out = mask_out[None]
        disp = out if out is not None else np.empty_like(L)
        mask = None
        if params.lr_check:
            mask = mask_out if mask_out is not None else np.empty_like(L)
        _same_shape(3, left=L, right=R, out=disp, mask_out=mask)
        _l.check(self._lib.gsm_stereo_batch(self._h, C.byref(params), n, _ptr(L), _ptr(R), _ptr(disp), _ptr(mask),
                                            rows, cols))
        if single and out is None:
            return disp[0], (mask[0] if mask is not None else None)
        return disp, mask

    def stereo_batch_async(self, left, right, params: GsmParams, out: np.ndarray, mask_out: Optional[np.ndarray] = None):
        """Streaming submit (gsm_stereo_batch_async): returns after enqueueing; results are valid after sync().
        left/right/out(/mask_out) must be C-contiguous uint8 [n, rows, cols], page-locked for real asynchrony, and
        must stay alive until sync()."""
        L, R = left, right
        n, rows, cols = _same_shape(3, left=L, right=R, out=out, mask_out=mask_out)
        _l.check(self._lib.gsm_stereo_batch_async(self._h, C.byref(params), n, _ptr(L), _ptr(R), _ptr(out),
                                                  _ptr(mask_out), rows, cols))

    def stereo(self, left, right, **kw):
        return self.stereo_batch(left, right, make_params(**kw))

    def stereo_batch_v(self, lefts, rights, params: GsmParams):
        """Mixed-size batch (gsm_stereo_batch_v): lists of [rows_i, cols_i] uint8 pairs -> (list of disparities, list of
        masks | None), the whole batch as ONE launch per stage.  BASELINE config 2 (nine Middlebury sets, three sizes);
        the separate Mats of Caller.cpp:12-19."""
        Ls = [_u8c(a, f"lefts[{i}]") for i, a in enumerate(lefts)]
        Rs = [_u8c(a, f"rights[{i}]") for i, a in enumerate(rights)]
        if len(Ls) != len(Rs) or not Ls:
            raise ValueError("lefts / rights must be non-empty lists of the same length")
        for i, (a, b) in enumerate(zip(Ls, Rs)):
            _same_shape(2, **{f"lefts[{i}]": a, f"rights[{i}]": b})
        n = len(Ls)
        disps = [np.empty_like(a) for a in Ls]
        masks = [np.empty_like(a) for a in Ls] if params.lr_check else None
        arr = lambda xs: (C.c_void_p * n)(*[x.ctypes.data for x in xs])
        rows = (C.c_int * n)(*[a.shape[0] for a in Ls])
        cols = (C.c_int * n)(*[a.shape[1] for a in Ls])
        _l.check(self._lib.gsm_stereo_batch_v(self._h, C.byref(params), n, arr(Ls), arr(Rs), arr(disps),
                                              arr(masks) if masks is not None else None, rows, cols))
        return disps, masks

    def stereo_device_v(self, left_ptr: int, right_ptr: int, disp_ptr: int, mask_ptr: int, shapes, params: GsmParams,
                        stream: int = 0) -> None:
        """Mixed-size batch on device buffers (gsm_stereo_device_v): shapes = [(rows_i, cols_i), ...], frames tight and
        concatenated in every buffer."""
        n = len(shapes)
        rows = (C.c_int * n)(*[int(s[0]) for s in shapes])
        cols = (C.c_int * n)(*[int(s[1]) for s in shapes])
        _l.check(self._lib.gsm_stereo_device_v(self._h, C.byref(params), n, C.c_void_p(left_ptr), C.c_void_p(right_ptr),
                                               C.c_void_p(disp_ptr), C.c_void_p(mask_ptr or None), rows, cols,
                                               C.c_void_p(stream or None)))

    def stereo_device(self, left_ptr: int, right_ptr: int, disp_ptr: int, mask_ptr: int, n: int, rows: int,
                      cols: int, params: GsmParams, stream: int = 0) -> None:
        """Device-resident buffers (raw device pointers, e.g. torch.Tensor.data_ptr()); asynchronous on `stream`."""
        _l.check(self._lib.gsm_stereo_device(self._h, C.byref(params), n, C.c_void_p(left_ptr), C.c_void_p(right_ptr),
                                             C.c_void_p(disp_ptr), C.c_void_p(mask_ptr or None), rows, cols,
                                             C.c_void_p(stream or None)))

    def sync(self):
        _l.check(self._lib.gsm_sync(self._h))

    # ---- multi-GPU disparity split -----------------------------------------------------------
    def partial_keys_device(self, left_ptr, right_ptr, keys_ptr, rows, cols, params: GsmParams, view=0, stream=0,
                            wait_event=0):
        """gsm_partial_keys_device; with wait_event (a cudaEvent_t handle, e.g. torch.cuda.Event.cuda_event) the fused
        kernel -- but not the disparity-independent passes in front of it -- waits for that event
        (gsm_partial_keys_device_ex)."""
        _l.check(self._lib.gsm_partial_keys_device_ex(self._h, C.byref(params), view, C.c_void_p(left_ptr),
                                                      C.c_void_p(right_ptr), C.c_void_p(keys_ptr), rows, cols,
                                                      C.c_void_p(stream or None), C.c_void_p(wait_event or None)))

    def finalize_keys_device(self, keys_left_ptr, keys_right_ptr, disp_ptr, mask_ptr, rows, cols, params: GsmParams,
                             stream=0):
        _l.check(self._lib.gsm_finalize_keys_device(self._h, C.byref(params), C.c_void_p(keys_left_ptr),
                                                    C.c_void_p(keys_right_ptr or None), C.c_void_p(disp_ptr),
                                                    C.c_void_p(mask_ptr or None), rows, cols,
                                                    C.c_void_p(stream or None)))

    def postfilter_device(self, disp_left_ptr, disp_right_ptr, disp_ptr, mask_ptr, rows, cols, params: GsmParams,
                          stream=0):
        """median on both views + LR check on device u8 maps (gsm_postfilter_device)."""
        _l.check(self._lib.gsm_postfilter_device(self._h, C.byref(params), C.c_void_p(disp_left_ptr),
                                                 C.c_void_p(disp_right_ptr or None), C.c_void_p(disp_ptr),
                                                 C.c_void_p(mask_ptr or None), rows, cols, C.c_void_p(stream or None)))

    def reduce_keys_p2p(self, key_ptrs, disp_ptrs, rank: int, npx: int, stream=0, max_blocks: int = 0):
        """Peer-memory combine of a disparity split (gsm_reduce_keys_p2p[_ex]): key_ptrs / disp_ptrs are the device
        pointers of every rank's packed-min plane / disparity map as mapped into this process.  max_blocks > 0 confines
        the kernel to that many SMs so that it runs beside a fused kernel that leaves them idle."""
        world = len(key_ptrs)
        ka = (C.c_void_p * world)(*[int(x) for x in key_ptrs])
        da = (C.c_void_p * world)(*[int(x) for x in disp_ptrs])
        _l.check(self._lib.gsm_reduce_keys_p2p_ex(self._h, ka, da, world, rank, npx, C.c_void_p(stream or None),
                                                  int(max_blocks)))

    def ad_volume(self, left, right, num_disp: int) -> np.ndarray:
        """== PreCal (BlockMatching.cpp:89-109): u8 [D][rows][cols]."""
        L, R = _u8c(left, "left"), _u8c(right, "right")
        _same_shape(2, left=L, right=R)
        out = np.empty((num_disp,) + L.shape, np.uint8)
        _l.check(self._lib.gsm_ad_volume(self._h, _ptr(L), _ptr(R), _ptr(out), L.shape[0], L.shape[1], num_disp))
        return out

    def cost_slices(self, left, right, params: GsmParams, d0: int, nd: int, view: int = 0) -> np.ndarray:
        """Aggregated cost for d in [d0, d0+nd): int32 SAD (mode sad) or float32 q (mode gf), [nd][rows][cols]."""
        L, R = _u8c(left, "left"), _u8c(right, "right")
        _same_shape(2, left=L, right=R)
        dt = np.int32 if params.mode == GSM_MODE_SAD else np.float32
        out = np.empty((nd,) + L.shape, dt)
        _l.check(self._lib.gsm_cost_slices(self._h, C.byref(params), view, _ptr(L), _ptr(R), d0, nd, _ptr(out),
                                           L.shape[0], L.shape[1]))
        return out

    def all_sad(self, left, right, radius: int, num_disp: int) -> np.ndarray:
        """== getAllSAD (BlockMatching.cpp:191-261): u8 [rows*cols][D]."""
        L, R = _u8c(left, "left"), _u8c(right, "right")
        _same_shape(2, left=L, right=R)
        out = np.empty((L.size, num_disp), np.uint8)
        _l.check(self._lib.gsm_all_sad(self._h, _ptr(L), _ptr(R), _ptr(out), L.shape[0], L.shape[1], radius,
                                       num_disp))
        return out

    def median(self, img, radius: int) -> np.ndarray:
        a = _u8c(img, "img")
        out = np.empty_like(a)
        _l.check(self._lib.gsm_median(self._h, _ptr(a), _ptr(out), a.shape[0], a.shape[1], radius))
        return out

    def lr_check(self, disp_left, disp_right):
        a, b = _u8c(disp_left, "disp_left"), _u8c(disp_right, "disp_right")
        _same_shape(2, disp_left=a, disp_right=b)
        occ, mask = np.empty_like(a), np.empty_like(a)
        _l.check(self._lib.gsm_lr_check(self._h, _ptr(a), _ptr(b), _ptr(occ), _ptr(mask), a.shape[0], a.shape[1]))
        return occ, mask

    # ---- the other two Device.cuh proxy functions ----------------------------------
    def remap(self, src, mapx, mapy) -> np.ndarray:
        """== kernalRemap / CPU_Remap: bilinear, OOB -> 0, round-nearest-even (Device.cu:127-167)."""
        a = _u8c(src, "src")
        mx = np.ascontiguousarray(mapx, np.float32); my = np.ascontiguousarray(mapy, np.float32)
        if a.ndim != 2 or mx.shape != a.shape or my.shape != a.shape:
            raise ValueError("src, mapx, mapy must be 2-D and the same size")
        out = np.empty_like(a)
        _l.check(self._lib.gsm_remap(self._h, _ptr(a), _ptr(mx), _ptr(my), _ptr(out), a.shape[0], a.shape[1]))
        return out

    def cvtcolor(self, src3, truncate: bool = False) -> np.ndarray:
        """== kernalCvtColor (round) / cvtColor_cpu (truncate): .299/.587/.114 on channels 0/1/2 as stored."""
        a = _u8c(src3, "src3")
        if a.ndim != 3 or a.shape[2] != 3:
            raise ValueError("src3 must be [rows, cols, 3]")
        out = np.empty(a.shape[:2], np.uint8)
        _l.check(self._lib.gsm_cvtcolor(self._h, _ptr(a), _ptr(out), a.shape[0], a.shape[1], int(truncate)))
        return out

    def disparity_to_depth(self, disp, fB: float) -> np.ndarray:
        """depth = fB / d (float32, 0 where d == 0): the Q-matrix depth of the rectified rig (Utility.cpp:228-234)."""
        a = _u8c(disp, "disp")
        if a.ndim != 2:
            raise ValueError("disp must be [rows, cols]")
        out = np.empty(a.shape, np.float32)
        _l.check(self._lib.gsm_disparity_to_depth(self._h, _ptr(a), _ptr(out), a.shape[0], a.shape[1], float(fB)))
        return out

    def set_rectification(self, mapx_left, mapy_left, mapx_right, mapy_right):
        """Upload the four CV_32FC1 maps Rectify() builds (Utility.cpp:228-234); params.rectify=True then feeds RAW
        frames, rectified on the device inside the plane packer.  Pass None four times to drop the maps."""
        if mapx_left is None:
            _l.check(self._lib.gsm_set_rectification(self._h, None, None, None, None, 0, 0))
            return
        maps = [np.ascontiguousarray(m, np.float32) for m in (mapx_left, mapy_left, mapx_right, mapy_right)]
        rows, cols = maps[0].shape
        if any(m.shape != (rows, cols) for m in maps):
            raise ValueError("the four maps must have the same shape")
        _l.check(self._lib.gsm_set_rectification(self._h, _ptr(maps[0]), _ptr(maps[1]), _ptr(maps[2]), _ptr(maps[3]),
                                                 rows, cols))

    # ---- segment-tree stereo (STMatching) -------------------------------------
    @staticmethod
    def _bgr(a, name):
        a = _u8c(a, name)
        if a.ndim != 3 or a.shape[2] != 3:
            raise ValueError(f"{name}: expected [rows, cols, 3] uint8 (BGR as cv::imread gives it)")
        return a

    def segment_tree_stereo(self, left_bgr, right_bgr, num_disp: int, sigma: float = 0.1, tau: float = 1200.0,
                            median_radius: int = 3, scale: int = 1, refined: bool = False) -> np.ndarray:
        """== stereo_disparity_normal (STMatching/StereoDisparity.cpp:58-90): colour+gradient cost -> segment-tree
        aggregation -> WTA -> median -> * scale; u8 [rows, cols].  refined=True: stereo_disparity_iteration (:92-160),
        the two-pass version with the L-R check and the colour+depth tree."""
        L, R = self._bgr(left_bgr, "left_bgr"), self._bgr(right_bgr, "right_bgr")
        if L.shape != R.shape:
            raise ValueError("left / right differ in shape")
        out = np.empty(L.shape[:2], np.uint8)
        p = GsmStParams(int(num_disp), float(sigma), float(tau), int(median_radius), int(scale), int(bool(refined)))
        _l.check(self._lib.gsm_segment_tree_stereo(self._h, C.byref(p), _ptr(L), _ptr(R), _ptr(out), L.shape[0], L.shape[1]))
        return out

    def segment_tree_stereo_batch(self, lefts_bgr, rights_bgr, num_disp: int, sigma: float = 0.1, tau: float = 1200.0,
                                  median_radius: int = 3, scale: int = 1, host_threads: int = 0) -> np.ndarray:
        """n pairs of one size ([n, rows, cols, 3] u8 each) through stereo_disparity_normal in one call: the trees of the
        frames are built concurrently on host threads (0 = one per hardware thread) while the GPU aggregates each frame as
        soon as its tree is ready; u8 [n, rows, cols], every map identical to segment_tree_stereo on that pair."""
        L, R = np.asarray(lefts_bgr), np.asarray(rights_bgr)
        for name, a in (("lefts_bgr", L), ("rights_bgr", R)):
            if a.dtype != np.uint8:
                raise TypeError(f"{name}: expected uint8, got {a.dtype}")
            if a.ndim != 4 or a.shape[3] != 3:
                raise ValueError(f"{name}: expected [n, rows, cols, 3], got {a.shape}")
        if L.shape != R.shape:
            raise ValueError("lefts / rights differ in shape")
        L, R = np.ascontiguousarray(L), np.ascontiguousarray(R)
        out = np.empty(L.shape[:3], np.uint8)
        p = GsmStParams(int(num_disp), float(sigma), float(tau), int(median_radius), int(scale), 0)
        _l.check(self._lib.gsm_segment_tree_stereo_batch(self._h, C.byref(p), _ptr(L), _ptr(R), _ptr(out), L.shape[0],
                                                         L.shape[1], L.shape[2], int(host_threads)))
        return out

    def st_matching_cost(self, left_bgr, right_bgr, num_disp: int) -> np.ndarray:
        """== GetMatchingCost (StereoHelper.cpp:75-129): float32 [rows, cols, num_disp]."""
        L, R = self._bgr(left_bgr, "left_bgr"), self._bgr(right_bgr, "right_bgr")
        if L.shape != R.shape:
            raise ValueError("left / right differ in shape")
        out = np.empty(L.shape[:2] + (num_disp,), np.float32)
        _l.check(self._lib.gsm_st_matching_cost(self._h, _ptr(L), _ptr(R), _ptr(out), L.shape[0], L.shape[1], num_disp))
        return out

    def st_filter(self, image_bgr, cost, sigma: float = 0.1, tau: float = 1200.0):
        """== CColorWeight + BuildSegmentTree + Filter (SegmentTree.cpp:38-195) on a float32 [rows, cols, D] volume:
        returns (aggregated volume, order, father_id, father_dist) -- the ordered tree in breadth-first order."""
        img = self._bgr(image_bgr, "image_bgr")
        vol = np.ascontiguousarray(cost, np.float32).copy()
        if vol.ndim != 3 or vol.shape[:2] != img.shape[:2]:
            raise ValueError("cost must be [rows, cols, D] for the image's rows, cols")
        n = img.shape[0] * img.shape[1]
        order, father = np.empty(n, np.int32), np.empty(n, np.int32)
        fdist = np.empty(n, np.uint8)
        _l.check(self._lib.gsm_st_filter(self._h, _ptr(img), _ptr(vol), img.shape[0], img.shape[1], vol.shape[2],
                                         float(sigma), float(tau), _ptr(order), _ptr(father), _ptr(fdist)))
        return vol, order, father, fdist

    # ---- introspection -----------------------------------------------------------------------
    @property
    def launch_count(self) -> int:
        return int(self._lib.gsm_launch_count(self._h))

    def set_kernel_timing(self, on: bool):
        _l.check(self._lib.gsm_set_kernel_timing(self._h, int(on)))

    def last_kernel_ms(self) -> float:
        return float(self._lib.gsm_last_kernel_ms(self._h))

    def measure_alu_peak(self) -> float:
        v = C.c_double(0.0)
        _l.check(self._lib.gsm_measure_alu_peak(self._h, C.byref(v)))
        return v.value


# ---- free functions with the reference's names ---------------------------------------------------
_default_ctx: Optional[StereoContext] = None


def _ctx_for(rows: int, cols: int, num_disp: int) -> StereoContext:
    global _default_ctx
    c = _default_ctx
    if c is None or rows > c.max_rows or cols > c.max_cols or num_disp > c.max_disp:
        if c is not None:
            c.close()
        _default_ctx = c = StereoContext(max(rows, 1080), max(cols, 1920), 256, 1)
    return c


def st_build_tree_host(wr, wu, tau: float = 1200.0):
    """Host stage of the segment-tree stereo (gsm_st_build_tree_host; no GPU): u8 edge weights [rows, cols] ->
    (order, father_id, father_dist, levels), the reference's ordered tree (CSegmentTree::m_tree)."""
    a, b = _u8c(wr, "wr"), _u8c(wu, "wu")
    _same_shape(2, wr=a, wu=b)
    n = a.size
    order, father = np.empty(n, np.int32), np.empty(n, np.int32)
    fdist = np.empty(n, np.uint8)
    levels = C.c_int(0)
    _l.check(_l.load().gsm_st_build_tree_host(_ptr(a), _ptr(b), a.shape[0], a.shape[1], float(tau), _ptr(order),
                                              _ptr(father), _ptr(fdist), C.byref(levels)))
    return order, father, fdist, levels.value


def blockMatching_gpu(h_left, h_right, SADWindowSize: int, searchRange: int) -> np.ndarray:
    """Drop-in for blockMatching_gpu(h_left, h_right, h_disparity, SADWindowSize, searchRange)
    (BlockMatching/Device.cuh:50): returns h_disparity."""
    L = _u8c(h_left, "h_left")
    return _ctx_for(L.shape[0], L.shape[1], searchRange).block_matching(L, h_right, SADWindowSize, searchRange)


def singleFrame(left_gray, right_gray) -> np.ndarray:
    """Compute part of singleFrame() (BlockMatching/Caller.cpp:9-25): blockMatching_gpu(g1, g2, disp, 5, 64).
    Image loading (imread + cvtColor) and display (imshow/waitKey) stay with the caller."""
    return blockMatching_gpu(left_gray, right_gray, 5, 64)


def remap_gpu(left, right, mapX1, mapY1, mapX2, mapY2) -> np.ndarray:
    """Drop-in for remap_gpu(left, right, mapX1, mapY1, mapX2, mapY2, rows, cols, total, result)
    (BlockMatching/Device.cuh:51, Device.cu:303-342): like the reference it returns only the LEFT remapped image
    (Device.cu:341); the right image is remapped too and discarded."""
    L = _u8c(left, "left")
    c = _ctx_for(L.shape[0], L.shape[1], 1)
    res = c.remap(L, mapX1, mapY1)
    c.remap(right, mapX2, mapY2)
    return res


def cvtColor_gpu(src3) -> np.ndarray:
    """Drop-in for cvtColor_gpu(uchar3* src, uchar* dst, rows, cols) (Device.cuh:52, Device.cu:344-367)."""
    a = _u8c(src3, "src3")
    return _ctx_for(a.shape[0], a.shape[1], 1).cvtcolor(a, truncate=False)


def compare_disp(reference_disp, gpu_disp):
    """compareDisp (BlockMatching.cpp:278-293) without the printing: list of (row, col, cpu, gpu) mismatches."""
    a, b = np.asarray(reference_disp), np.asarray(gpu_disp)
    ys, xs = np.nonzero(a != b)
    return [(int(y), int(x), int(a[y, x]), int(b[y, x])) for y, x in zip(ys, xs)]
