"""Segment-tree stereo (gsm_segment_tree_stereo) timing on one GPU: whole call from host buffers, and the host tree
builder alone (dev tool).  usage: python tools/st_bench.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.ndimage import median_filter
import gpu_stereo_matching_b200 as g
from gpu_stereo_matching_b200 import data


def weights(img):
    med = np.stack([median_filter(img[:, :, c], size=3, mode="nearest") for c in range(3)], -1).astype(np.int32)
    wr = np.full(img.shape[:2], 255, np.uint8); wu = np.full(img.shape[:2], 255, np.uint8)
    wr[:, :-1] = np.abs(med[:, :-1] - med[:, 1:]).max(-1); wu[1:, :] = np.abs(med[1:, :] - med[:-1, :]).max(-1)
    return wr, wu


for (h, w, D) in ((256, 320, 64), (370, 463, 64), (720, 1280, 128)):
    L, R = data.synthetic_color_pair(h, w, 7, dmax=min(D - 8, 120))
    with g.StereoContext(h, w, D, 1) as ctx:
        for _ in range(2):
            d = ctx.segment_tree_stereo(L, R, D, scale=1)
        t0 = time.perf_counter()
        for _ in range(5):
            ctx.segment_tree_stereo(L, R, D, scale=1)
        ms = (time.perf_counter() - t0) / 5 * 1e3
        ms2 = None
        if D <= w and h * w < 500000:  # ST-2 (stereo_disparity_iteration): three trees, both views
            ctx.segment_tree_stereo(L, R, D, scale=1, refined=True)
            t0 = time.perf_counter()
            for _ in range(3):
                ctx.segment_tree_stereo(L, R, D, scale=1, refined=True)
            ms2 = (time.perf_counter() - t0) / 3 * 1e3
        nb = 32 if h * w < 500000 else 8  # a batch: the trees of the frames are built on all host threads
        Lb, Rb = np.stack([L] * nb), np.stack([R] * nb)
        ctx.segment_tree_stereo_batch(Lb, Rb, D)
        t0 = time.perf_counter()
        bd = ctx.segment_tree_stereo_batch(Lb, Rb, D)
        bms = (time.perf_counter() - t0) * 1e3 / nb
        assert all(np.array_equal(bd[i], d) for i in range(nb))
    wr, wu = weights(L)
    t0 = time.perf_counter()
    for _ in range(5):
        _, _, _, levels = g.st_build_tree_host(wr, wu)
    tms = (time.perf_counter() - t0) / 5 * 1e3
    print(f"{w}x{h} x{D}: whole call {ms:8.2f} ms ({h * w * D / ms / 1e3:8.1f} MDE/s), host tree builder {tms:7.2f} ms, "
          f"tree depth {levels} levels" + (f", ST-2 whole call {ms2:8.2f} ms" if ms2 else "")
          + f", batch of {nb}: {bms:6.2f} ms per pair ({1e3 / bms:6.1f} pairs/s)", flush=True)
