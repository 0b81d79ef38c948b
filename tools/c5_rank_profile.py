"""One rank's share of config 5 at a given world size (d in [0, 256/world) of a 3840x2160 x256 pair, the row bands the
split uses) on ONE GPU: time of the partial-keys step and of its fused kernel (dev tool; run it under
`ncu --metrics gpu__time_duration.sum` for the per-kernel list).  usage: python tools/c5_rank_profile.py [world=8] [rank=0]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpu_stereo_matching_b200 as g
from gpu_stereo_matching_b200 import data
from gpu_stereo_matching_b200.dist import torch_stream_handle, dsplit_row_bands, shard_disparities
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
h, w, d = 2160, 3840, 256
L, R, _ = data.synthetic_pair(h, w, 3000, dmax=250)
Ld, Rd = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
kl = torch.empty(h * w, dtype=torch.int64, device="cuda")
ctx = g.StereoContext(h, w, d, 1)
st = torch.cuda.Stream(); sh = torch_stream_handle(st)
bands = dsplit_row_bands(h, w, d, world)
d0, d1 = shard_disparities(d, world, rank)
pp = g.make_params("gf", 9, d, row_bands=bands, d_begin=d0, d_end=d1)
ctx.set_kernel_timing(True)
with torch.cuda.stream(st):
    for _ in range(3):
        ctx.partial_keys_device(Ld.data_ptr(), Rd.data_ptr(), kl.data_ptr(), h, w, pp, 0, sh)
    st.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(st)
        ctx.partial_keys_device(Ld.data_ptr(), Rd.data_ptr(), kl.data_ptr(), h, w, pp, 0, sh)
        e1.record(st); st.synchronize()
        ts.append((e0.elapsed_time(e1), ctx.last_kernel_ms()))
t = min(ts)
print(f"world {world}: d [{d0},{d1}) bands {bands}: partial keys {t[0]:.3f} ms, fused kernel {t[1]:.3f} ms, rest {t[0] - t[1]:.3f} ms", flush=True)
