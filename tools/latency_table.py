"""GF latency vs batch size at 720p x 128 d (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpu_stereo_matching_b200 as g
from gpu_stereo_matching_b200 import data
from gpu_stereo_matching_b200.dist import torch_stream_handle
L, R = data.synthetic_batch(2, 720, 1280, 1234)
ctx = g.StereoContext(720, 1280, 128, 32)
st = torch.cuda.Stream(); sh = torch_stream_handle(st)
for n in (1, 2, 4, 8, 16, 32):
    Ld = torch.from_numpy(np.tile(L, (max(1, n // 2) + 1, 1, 1))[:n]).cuda(); Rd = torch.from_numpy(np.tile(R, (max(1, n // 2) + 1, 1, 1))[:n]).cuda()
    Dd = torch.empty_like(Ld)
    for bands in (0, 1, 2, 3, 4, 6):
        p = g.make_params("gf", 9, 128, row_bands=bands)
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
        print(f"n={n:2d} bands={bands}: {ms:7.3f} ms/batch  {ms/n:6.3f} ms/frame  {n/ms*1e3:7.0f} fps", flush=True)
