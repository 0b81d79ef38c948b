"""SAD-mode micro-bench (dev tool): 720p x 128 d, r = 5, device-resident batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpu_stereo_matching_b200 as g
from gpu_stereo_matching_b200 import data
from gpu_stereo_matching_b200.dist import torch_stream_handle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
r = int(sys.argv[2]) if len(sys.argv) > 2 else 5
L, R = data.synthetic_batch(2, 720, 1280, 1234)
Ld = torch.from_numpy(np.tile(L, (n // 2, 1, 1))).cuda(); Rd = torch.from_numpy(np.tile(R, (n // 2, 1, 1))).cuda()
Dd = torch.empty_like(Ld)
ctx = g.StereoContext(720, 1280, 128, n)
p = g.make_params("sad", r, 128)
st = torch.cuda.Stream(); sh = torch_stream_handle(st)
ctx.set_kernel_timing(True)
with torch.cuda.stream(st):
    for _ in range(3):
        ctx.stereo_device(Ld.data_ptr(), Rd.data_ptr(), Dd.data_ptr(), 0, n, 720, 1280, p, sh)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(st)
    for _ in range(5):
        ctx.stereo_device(Ld.data_ptr(), Rd.data_ptr(), Dd.data_ptr(), 0, n, 720, 1280, p, sh)
    e1.record(st)
st.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"SAD r={r} n={n}: {ms:.3f} ms/batch, {n/ms*1e3:.0f} fps, {n*720*1280*128/ms/1e6:.0f} GDE/s, fused kernel {ctx.last_kernel_ms():.3f} ms")
