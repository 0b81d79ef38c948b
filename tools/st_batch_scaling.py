"""gsm_segment_tree_stereo_batch: ms per pair against the number of builder threads and the batch size (dev tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpu_stereo_matching_b200 as g
from gpu_stereo_matching_b200 import data

h, w, D = 370, 463, 64
L, R = data.synthetic_color_pair(h, w, 7, dmax=D - 8)
with g.StereoContext(h, w, D, 1) as ctx:
    for nb in (8, 32, 64):
        Lb, Rb = np.stack([L] * nb), np.stack([R] * nb)
        for th in (1, 2, 4, 8, 12, 16):
            ctx.segment_tree_stereo_batch(Lb, Rb, D, host_threads=th)
            t0 = time.perf_counter()
            for _ in range(2):
                ctx.segment_tree_stereo_batch(Lb, Rb, D, host_threads=th)
            ms = (time.perf_counter() - t0) * 1e3 / 2 / nb
            print(f"{w}x{h} x{D}: batch {nb:3d}, {th:2d} builder threads: {ms:6.2f} ms per pair ({1e3 / ms:7.1f} pairs/s)", flush=True)
