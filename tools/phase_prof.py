"""Per-phase cycle shares of gf3_wta_kernel (dev tool).

Needs the instrumented build of the library:
    make -C gpu_stereo_matching_b200/csrc OUT=../../tools/variants/libgsm_prof.so EXTRA=-DGSM_GF_PROFILE
    python tools/phase_prof.py tools/variants/libgsm_prof.so
GSM_GF_PROFILE adds clock() reads between the phases of a march step and a debug export; the product build has
neither (identical SASS with the macro off)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpu_stereo_matching_b200 import lib
lib.LIB_PATH = os.path.abspath(sys.argv[1])
import gpu_stereo_matching_b200 as g
from gpu_stereo_matching_b200 import data
from gpu_stereo_matching_b200.dist import torch_stream_handle
n = 16
L, R = data.synthetic_batch(4, 720, 1280, 1234)
Ld = torch.from_numpy(np.tile(L, (n // 4, 1, 1))).cuda(); Rd = torch.from_numpy(np.tile(R, (n // 4, 1, 1))).cuda()
Dd = torch.empty_like(Ld)
ctx = g.StereoContext(720, 1280, 128, n)
st = torch.cuda.Stream(); sh = torch_stream_handle(st)
p = g.make_params("gf", 9, 128)
l = lib.load()
buf = (C.c_ulonglong * (16 * 9))()
with torch.cuda.stream(st):
    ctx.stereo_device(Ld.data_ptr(), Rd.data_ptr(), Dd.data_ptr(), 0, n, 720, 1280, p, sh); st.synchronize()
    l.gsm_debug_gf_prof(None, 1)
    ctx.stereo_device(Ld.data_ptr(), Rd.data_ptr(), Dd.data_ptr(), 0, n, 720, 1280, p, sh); st.synchronize()
    l.gsm_debug_gf_prof(buf, 0)
a = np.array(list(buf), dtype=np.float64).reshape(16, 9)[:12]
names = ["mbar wait", "AD+init", "slide", "ab", "store", "barrier+issue", "B loads+slide", "q", "WTA"]
tot = a.sum(axis=1, keepdims=True)
print("run  " + " ".join(f"{x:>14s}" for x in names) + "   total(Mcyc)")
for r in range(12):
    print(f"{r:3d}  " + " ".join(f"{100 * a[r, i] / tot[r, 0]:13.1f}%" for i in range(9)) + f"   {tot[r, 0] / 1e6:9.1f}")
m = a[1:11].sum(axis=0); print("int  " + " ".join(f"{100 * x / m.sum():13.1f}%" for x in m))
