import sys, numpy as np
sys.path.insert(0,'.')
import gpu_stereo_matching_b200 as g
from oracle import oracle as O
fx = np.load("tests/golden/middlebury_gray.npz")
ctx = g.StereoContext(1080,1920,256,2)
h,w=60,200
stripes = np.tile((np.arange(w)//10%2*255).astype(np.uint8),(h,1))
cases = [("stripes", stripes, np.zeros((h,w),np.uint8), 9, 48)] + [(n, fx[n+"_L"], fx[n+"_R"], 9, 64) for n in ("Art","Reindeer","Books")] + [("Books5", fx["Books_L"], fx["Books_R"], 5, 32)]
for name,L,R,r,D in cases:
    p = g.make_params("gf", r, D)
    q = ctx.cost_slices(L,R,p,0,D); qr = O.gf_cost_slices(L,R,r,0,D)
    err = np.abs(q-qr)/np.maximum(np.abs(qr),1)
    print(name, r, "max", err.max(), "mean", err.mean())
