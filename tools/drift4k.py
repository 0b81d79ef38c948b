import sys, numpy as np
sys.path.insert(0,'.')
import gpu_stereo_matching_b200 as g
from gpu_stereo_matching_b200 import data
from oracle import oracle as O
L,R,_ = data.synthetic_pair(2160,3840,3000,dmax=250)
ctx = g.StereoContext(2160,3840,256,1)
p = g.make_params("gf",9,256)
for bands in (1,0):
    p.row_bands = bands
    q = ctx.cost_slices(L,R,p,100,4)
    qr = O.gf_cost_slices(L,R,9,100,4)
    err = np.abs(q-qr)/np.maximum(np.abs(qr),1)
    print("bands",bands,"4K max rel err", err.max(), "rows of max", np.unravel_index(err.argmax(), err.shape), "mean", err.mean())
    # error vs row (drift?)
    print("  max err by row block:", [float(err[:, i:i+270].max()) for i in range(0,2160,270)])
