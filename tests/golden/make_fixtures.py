"""Regenerates tests/golden/*.npz|json from /root/reference (dev container only).

  middlebury_gray.npz : gray u8 left/right for the 9 Middlebury-2005 third-size sets shipped in
                        /root/reference/Images (view1 = left, view5 = right) plus the 320x256
                        Art pair the reference demo actually runs (Caller.cpp:12-13), converted
                        with cv2.cvtColor(cv2.imread(p), COLOR_BGR2GRAY) like Caller.cpp:15-16.
                        These are DATA (public Middlebury images), not reference sources.
  ref_digests.json    : outputs of the reference's OWN CPU code compiled unmodified
                        (oracle/_ref/libref.so): getDisp digests (FNV-1a-64, sum, zero count) for
                        every pair at r=5, D=64 (the Caller.cpp:19 parameters), plus
                        PreCal / getAllSAD / ctmf digests on the Art demo pair.
Run:  python tests/golden/make_fixtures.py
"""
import json
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as O  # noqa: E402

REF = "/root/reference/Images"
SETS = ["Art", "Books", "Computer", "Dolls", "Drumsticks", "Dwarves", "Laundry", "Moebius", "Reindeer"]


def gray(p):
    img = cv2.imread(p)
    assert img is not None, p
    return cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)


def main():
    arrays = {}
    for s in SETS:
        arrays[s + "_L"] = gray(f"{REF}/{s}/view1.png")
        arrays[s + "_R"] = gray(f"{REF}/{s}/view5.png")
    arrays["ArtDemo_L"] = gray(f"{REF}/Art/view1_.png")
    arrays["ArtDemo_R"] = gray(f"{REF}/Art/view5_.png")
    np.savez_compressed(os.path.join(HERE, "middlebury_gray.npz"), **arrays)

    dig = {"params": {"radius": 5, "D": 64}, "getDisp": {}}
    for s in SETS + ["ArtDemo"]:
        L, R = arrays[s + "_L"], arrays[s + "_R"]
        d = O.ref_getDisp(L, R, 5, 64)
        dig["getDisp"][s] = {"shape": list(L.shape), "fnv1a64": O.fnv1a64(d), "sum": int(d.sum()),
                             "zeros": int((d == 0).sum())}
        print(s, dig["getDisp"][s])
    L, R = arrays["ArtDemo_L"], arrays["ArtDemo_R"]
    dig["PreCal_ArtDemo_D64"] = O.fnv1a64(O.ref_PreCal(L, R, 64))
    dig["getAllSAD_ArtDemo_r5_D64"] = O.fnv1a64(O.ref_getAllSAD(L, R, 5, 64))
    d = O.ref_getDisp(L, R, 5, 64)
    dig["ctmf_r3_on_getDisp_ArtDemo"] = O.fnv1a64(O.ref_median(d, 3))
    dig["getDisp_ArtDemo_r9_D64"] = O.fnv1a64(O.ref_getDisp(L, R, 9, 64))
    dig["getDisp_ArtDemo_r2_D16"] = O.fnv1a64(O.ref_getDisp(L, R, 2, 16))
    with open(os.path.join(HERE, "ref_digests.json"), "w") as f:
        json.dump(dig, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
