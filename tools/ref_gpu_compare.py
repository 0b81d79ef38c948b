"""The reference's OWN CUDA code (BlockMatching/Device.cu, compiled unmodified for sm_100a into
oracle/_ref/libdevref.so by `make -C oracle devref`) run on the same GPU next to libgsm.so (dev tool, not a test).

Reports, for the reference's demo input (Art 320x256, SAD radius 5, 64 disparities -- the only size its launch
geometry covers): disparity mismatches between blockMatching_gpu (reference), getDisp (reference CPU, via the
oracle) and gsm_block_matching (this repo), and wall time per call; likewise remap_gpu / cvtColor_gpu."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import gpu_stereo_matching_b200 as g
from oracle import oracle as O

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
dev = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libdevref.so"))
dev.devref_block_matching.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, u8p]
dev.devref_remap.argtypes = [u8p, u8p, f32p, f32p, f32p, f32p, C.c_int, C.c_int, u8p]
dev.devref_cvtcolor.argtypes = [u8p, u8p, C.c_int, C.c_int]

fx = np.load(os.path.join(ROOT, "tests", "golden", "middlebury_gray.npz"))
L, R = np.ascontiguousarray(fx["ArtDemo_L"]), np.ascontiguousarray(fx["ArtDemo_R"])
h, w = L.shape
print(f"input: Art demo pair {w}x{h}, SAD radius 5, 64 disparities", flush=True)
ref_gpu = np.empty_like(L)
with O.quiet_stdout():                      # the reference prints its phase timings
    for _ in range(3):
        rc = dev.devref_block_matching(L, R, h, w, 5, 64, ref_gpu)
    t0 = time.perf_counter()
    for _ in range(5):
        rc = dev.devref_block_matching(L, R, h, w, 5, 64, ref_gpu)
    t_ref = (time.perf_counter() - t0) / 5
print(f"reference blockMatching_gpu on this GPU: rc={rc}, {t_ref * 1e3:.3f} ms per call (host buffers in, host buffer out)")
ctx = g.StereoContext(h, w, 64, 1)
for _ in range(3):
    ours = ctx.block_matching(L, R, 5, 64)
t0 = time.perf_counter()
for _ in range(20):
    ours = ctx.block_matching(L, R, 5, 64)
t_ours = (time.perf_counter() - t0) / 20
print(f"gsm_block_matching (this repo):          {t_ours * 1e3:.3f} ms per call (same buffers)  -> {t_ref / t_ours:.1f}x")
cpu = O.ref_getDisp(L, R, 5, 64) if O.have_ref() else O.sad_wta(L, R, 5, 64)
print("mismatching pixels: reference GPU vs reference CPU getDisp:", int((ref_gpu != cpu).sum()),
      "| this repo vs reference GPU:", int((ours != ref_gpu).sum()), "| this repo vs reference CPU:", int((ours != cpu).sum()))

rng = np.random.default_rng(3)
hh, ww = 200, 320
img = rng.integers(0, 256, (hh, ww), dtype=np.uint8)
mx = (rng.random((hh, ww), dtype=np.float32) * (ww + 6) - 3).astype(np.float32)
my = (rng.random((hh, ww), dtype=np.float32) * (hh + 6) - 3).astype(np.float32)
res = np.empty_like(img)
with O.quiet_stdout():
    rc = dev.devref_remap(img, img, mx, my, mx, my, hh, ww, res)
mine = ctx2 = None
ctx2 = g.StereoContext(hh, ww, 64, 1)
mine = ctx2.remap(img, mx, my)
cpu_twin = O.remap(img, mx, my)
d = np.abs(res.astype(int) - mine.astype(int))
print(f"remap {ww}x{hh}: reference kernalRemap vs gsm_remap: {int((d != 0).sum())} pixels differ (max {int(d.max())});"
      f" reference kernalRemap vs its CPU twin CPU_Remap: {int((res != cpu_twin).sum())}; gsm_remap vs CPU_Remap: {int((mine != cpu_twin).sum())}")
rgb = rng.integers(0, 256, (hh, ww, 3), dtype=np.uint8)
gray = np.empty((hh, ww), np.uint8)
with O.quiet_stdout():
    rc = dev.devref_cvtcolor(rgb.reshape(-1), gray, hh, ww)
print(f"cvtColor {ww}x{hh}: reference kernalCvtColor vs gsm_cvtcolor (rounding): {int((gray != ctx2.cvtcolor(rgb)).sum())} pixels differ;"
      f" vs the rounding restatement of the oracle: {int((gray != O.cvtcolor(rgb, truncate=False)).sum())}")
