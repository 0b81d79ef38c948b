import sys, numpy as np
sys.path.insert(0,'.')
import gpu_stereo_matching_b200 as g
from gpu_stereo_matching_b200 import data
from oracle import oracle as O
fx = np.load("tests/golden/middlebury_gray.npz")
ctx = g.StereoContext(1080,1920,256,2)
Ls, Rs, _ = data.synthetic_pair(72, 420, 903, dmax=200)
for r in range(1, 10):
    worst = 0
    for name,(L,R,D) in {"Laundry":(fx["Laundry_L"],fx["Laundry_R"],40), "Books":(fx["Books_L"],fx["Books_R"],32), "synth":(Ls,Rs,64)}.items():
        for view in (0,1):
            p = g.make_params("gf", r, D)
            q = ctx.cost_slices(L,R,p,0,D,view=view); qr = O.gf_cost_slices(L,R,r,0,D,view=view)
            worst = max(worst, float((np.abs(q-qr)/np.maximum(np.abs(qr),1)).max()))
    print("r", r, "worst", worst)
