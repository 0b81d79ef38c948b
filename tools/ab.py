"""A/B timing of libgsm.so build variants (dev tool).  usage: python tools/ab.py [variant.so ...]
Each variant runs in a fresh process: GF r=9, 32 frames 1280x720 x128d resident in HBM, fused-kernel ms and checksum."""
import os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import numpy as np, torch, zlib
    from gpu_stereo_matching_b200 import lib
    if sys.argv[2] != "default":
        lib.LIB_PATH = os.path.abspath(sys.argv[2])
    import gpu_stereo_matching_b200 as g
    from gpu_stereo_matching_b200 import data
    from gpu_stereo_matching_b200.dist import torch_stream_handle
    mode, r, D, n = os.environ.get("AB_MODE", "gf"), int(os.environ.get("AB_R", 9)), int(os.environ.get("AB_D", 128)), int(os.environ.get("AB_N", 32))
    H, W = int(os.environ.get("AB_H", 720)), int(os.environ.get("AB_W", 1280))
    L, R = data.synthetic_batch(min(4, n), H, W, 1234, dmax=min(D - 8, 120))
    Ld = torch.from_numpy(np.tile(L, (n // len(L) + 1, 1, 1))[:n]).cuda(); Rd = torch.from_numpy(np.tile(R, (n // len(R) + 1, 1, 1))[:n]).cuda()
    Dd = torch.empty_like(Ld)
    ctx = g.StereoContext(H, W, D, n)
    st = torch.cuda.Stream(); sh = torch_stream_handle(st)
    p = g.make_params(mode, r, D)
    ctx.set_kernel_timing(True)
    with torch.cuda.stream(st):
        for _ in range(3):
            ctx.stereo_device(Ld.data_ptr(), Rd.data_ptr(), Dd.data_ptr(), 0, n, H, W, p, sh)
        st.synchronize()
        best = 1e9; kbest = 1e9
        for _ in range(6):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(st)
            ctx.stereo_device(Ld.data_ptr(), Rd.data_ptr(), Dd.data_ptr(), 0, n, H, W, p, sh)
            e1.record(st); st.synchronize()
            best = min(best, e0.elapsed_time(e1)); kbest = min(kbest, ctx.last_kernel_ms())
    crc = zlib.crc32(Dd.cpu().numpy().tobytes())
    print(f"{os.path.basename(sys.argv[2]):28s} {mode} {W}x{H} r={r} D={D} n={n}: step {best:7.3f} ms  fused {kbest:7.3f} ms  {n / best * 1e3:7.0f} fps  crc {crc:08x}", flush=True)
else:
    for v in (sys.argv[1:] or ["default"]):
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", v], check=False)
