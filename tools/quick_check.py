"""Scratch GPU check: SAD bit-exactness + GF error statistics vs the oracle (dev tool, not a test)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpu_stereo_matching_b200 as g
from oracle import oracle as O

fx = np.load("tests/golden/middlebury_gray.npz")
ctx = g.StereoContext(1080, 1920, 256, 4)
ok = True
for name, r, D in [("ArtDemo", 5, 64), ("Art", 5, 64), ("Art", 9, 64), ("Laundry", 3, 48), ("Art", 1, 16), ("Art", 12, 100)]:
    L, R = fx[name + "_L"], fx[name + "_R"]
    ref = O.sad_wta(L, R, r, D)
    t = time.time(); out = ctx.block_matching(L, R, r, D); dt = time.time() - t
    mm = int((ref != out).sum())
    print(f"SAD {name} r={r} D={D}: mismatches {mm} ({dt*1e3:.2f} ms)")
    ok &= mm == 0
L, R = fx["ArtDemo_L"], fx["ArtDemo_R"]
ad = ctx.ad_volume(L, R, 64); print("AD volume mismatch", int((ad != O.ad_volume(L, R, 64)).sum()))
p = g.make_params("sad", 5, 64)
cs = ctx.cost_slices(L, R, p, 30, 5)
refs = np.stack([O.sad_slice(L, R, 5, d) for d in range(30, 35)])
print("SAD slices mismatch", int((cs != refs).sum()))
alls = ctx.all_sad(L, R, 5, 64); print("allSAD mismatch", int((alls != O.all_sad(L, R, 5, 64)).sum()))
d = ctx.block_matching(L, R, 5, 64)
print("median mismatch", int((ctx.median(d, 3) != O.median(d, 3)).sum()))
# GF
for name, r, D in [("ArtDemo", 9, 64), ("Art", 9, 64), ("Art", 4, 32)]:
    L, R = fx[name + "_L"], fx[name + "_R"]
    p = g.make_params("gf", r, D)
    for view in (0, 1):
        q = ctx.cost_slices(L, R, p, 0, D, view=view).astype(np.float64)
        qr = O.gf_cost_slices(L, R, r, 0, D, view=view)
        err = np.abs(q - qr) / np.maximum(np.abs(qr), 1.0)
        print(f"GF {name} r={r} D={D} view={view}: max rel err {err.max():.3e} mean {err.mean():.3e} 99.99pct {np.quantile(err, 0.9999):.3e} q range [{qr.min():.2f},{qr.max():.2f}]")
    disp, _ = ctx.stereo_batch(L, R, p)
    dref = O.gf_wta(L, R, r, D)
    diff = np.abs(disp.astype(int) - dref.astype(int))
    print(f"   WTA identical {100*(diff==0).mean():.4f}%  >1: {int((diff>1).sum())}")
    p2 = g.make_params("gf", r, D, lr_check=True, median_radius=3)
    disp2, mask2 = ctx.stereo_batch(L, R, p2)
    dref2, mref2 = O.stereo_pipeline(L, R, mode="gf", r=r, D=D, lr=True, median_r=3)
    diff2 = np.abs(disp2.astype(int) - dref2.astype(int))
    print(f"   LR+median identical {100*(diff2==0).mean():.4f}%  >1: {int((diff2>1).sum())} mask identical {100*(mask2==mref2).mean():.4f}%")
# timing at 720p x 128
import torch
rng = np.random.default_rng(0)
n = 4
Lb = rng.integers(0, 256, (n, 720, 1280), dtype=np.uint8); Rb = np.roll(Lb, -7, axis=2)
Ld, Rd = torch.from_numpy(Lb).cuda(), torch.from_numpy(Rb).cuda()
Dd = torch.empty_like(Ld)
for mode, r in (("sad", 5), ("sad", 9), ("gf", 9)):
    p = g.make_params(mode, r, 128)
    ctx.set_kernel_timing(True)
    for _ in range(3):
        ctx.stereo_device(Ld.data_ptr(), Rd.data_ptr(), Dd.data_ptr(), 0, n, 720, 1280, p, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        ctx.stereo_device(Ld.data_ptr(), Rd.data_ptr(), Dd.data_ptr(), 0, n, 720, 1280, p, torch.cuda.current_stream().cuda_stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    kms = ctx.last_kernel_ms()
    de = n * 720 * 1280 * 128
    print(f"{mode} r={r} 720p x128 batch {n}: {ms:.3f} ms/batch -> {n/ms*1e3:.1f} fps, {de/ms/1e6:.1f} GDE/s; fused kernel {kms:.3f} ms")
print("alu peak Tlaneop/s", ctx.measure_alu_peak() / 1e12)
print("ALL SAD OK" if ok else "SAD MISMATCH")
