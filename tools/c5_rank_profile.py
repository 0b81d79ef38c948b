"""One rank's share of config 5 at world size 8 (d in [0, 32) of a 3840x2160 x256 pair) on one GPU (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpu_stereo_matching_b200 as g
from gpu_stereo_matching_b200 import data
from gpu_stereo_matching_b200.dist import torch_stream_handle
h, w, d = 2160, 3840, 256
L, R, _ = data.synthetic_pair(h, w, 3000, dmax=250)
Ld, Rd = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
Dd = torch.empty_like(Ld)
kl = torch.empty(h * w, dtype=torch.int64, device="cuda")
ctx = g.StereoContext(h, w, d, 1)
st = torch.cuda.Stream(); sh = torch_stream_handle(st)
p = g.make_params("gf", 9, d)
for nd in (32, 64, 128, 256):
    pp = g.make_params("gf", 9, d, d_begin=0, d_end=nd)
    ctx.set_kernel_timing(True)
    with torch.cuda.stream(st):
        for _ in range(3):
            ctx.partial_keys_device(Ld.data_ptr(), Rd.data_ptr(), kl.data_ptr(), h, w, pp, 0, sh)
            ctx.finalize_keys_device(kl.data_ptr(), 0, Dd.data_ptr(), 0, h, w, p, sh)
        st.synchronize()
        e0, e1, e2 = torch.cuda.Event(True), torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(st)
        ctx.partial_keys_device(Ld.data_ptr(), Rd.data_ptr(), kl.data_ptr(), h, w, pp, 0, sh)
        e1.record(st)
        ctx.finalize_keys_device(kl.data_ptr(), 0, Dd.data_ptr(), 0, h, w, p, sh)
        e2.record(st); st.synchronize()
    print(f"d range {nd:3d}: partial {e0.elapsed_time(e1):.3f} ms (fused kernel {ctx.last_kernel_ms():.3f}), finalize {e1.elapsed_time(e2):.3f} ms", flush=True)
