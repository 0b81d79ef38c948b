"""CPU tests: the oracle against the reference's golden outputs and against itself (no GPU)."""
import os

import numpy as np
import pytest

from conftest import ROOT, SETS


def test_getdisp_digests_all_sets(fx, digests, orc):
    """oracle.sad_wta == the reference's getDisp (compiled unmodified) on every shipped pair, r=5, D=64."""
    for s in SETS + ["ArtDemo"]:
        d = orc.sad_wta(fx[s + "_L"], fx[s + "_R"], 5, 64)
        g = digests["getDisp"][s]
        assert list(d.shape) == g["shape"]
        assert int(d.sum()) == g["sum"] and int((d == 0).sum()) == g["zeros"], s
        assert orc.fnv1a64(d) == g["fnv1a64"], s


def test_survey_probe_values(fx, orc):
    """Sum / zero counts recorded by the survey's probe of the compiled reference."""
    d = orc.sad_wta(fx["ArtDemo_L"], fx["ArtDemo_R"], 5, 64)
    assert (int(d.sum()), int((d == 0).sum())) == (2489456, 1702)
    d = orc.sad_wta(fx["Art_L"], fx["Art_R"], 5, 64)
    assert (int(d.sum()), int((d == 0).sum())) == (6887115, 3738)


def test_other_digests(fx, digests, orc):
    L, R = fx["ArtDemo_L"], fx["ArtDemo_R"]
    assert orc.fnv1a64(orc.ad_volume(L, R, 64)) == digests["PreCal_ArtDemo_D64"]
    assert orc.fnv1a64(orc.all_sad(L, R, 5, 64)) == digests["getAllSAD_ArtDemo_r5_D64"]
    d = orc.sad_wta(L, R, 5, 64)
    assert orc.fnv1a64(orc.median(d, 3)) == digests["ctmf_r3_on_getDisp_ArtDemo"]
    assert orc.fnv1a64(orc.sad_wta(L, R, 9, 64)) == digests["getDisp_ArtDemo_r9_D64"]
    assert orc.fnv1a64(orc.sad_wta(L, R, 2, 16)) == digests["getDisp_ArtDemo_r2_D16"]


def test_direct_loop_equals_box_sums(orc):
    """The literal getDisp loop structure and the O(1) box-sum version agree, incl. degenerate shapes."""
    rng = np.random.default_rng(7)
    for (h, w, r, D) in [(23, 31, 2, 12), (9, 40, 5, 33), (40, 9, 3, 16), (5, 5, 4, 8), (1, 17, 1, 4), (17, 1, 2, 3)]:
        L = rng.integers(0, 256, (h, w), dtype=np.uint8)
        R = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(orc.sad_wta(L, R, r, D), orc.sad_wta(L, R, r, D, direct=True)), (h, w, r, D)


def test_against_compiled_reference(orc):
    """Direct comparison with oracle/_ref/libref.so when it travelled with the repo."""
    if not orc.have_ref():
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference)")
    rng = np.random.default_rng(11)
    for (h, w, r, D) in [(37, 53, 5, 20), (30, 70, 2, 64), (64, 48, 7, 40)]:
        L = rng.integers(0, 256, (h, w), dtype=np.uint8)
        # correlated right image so that some SADs pass the 50*(2r+1)^2 threshold
        R = np.roll(L, -3, axis=1) ^ rng.integers(0, 8, (h, w), dtype=np.uint8)
        assert np.array_equal(orc.sad_wta(L, R, r, D), orc.ref_getDisp(L, R, r, D))
        assert np.array_equal(orc.ad_volume(L, R, D), orc.ref_PreCal(L, R, D))
        assert np.array_equal(orc.all_sad(L, R, r, D), orc.ref_getAllSAD(L, R, r, D))
    # MeanFilter passes memsize = area*channels (Toolkit.cpp:39-41): ctmf stripes need area/544 > 2r+1
    img = rng.integers(0, 64, (128, 160), dtype=np.uint8)
    for r in (1, 2, 3):
        assert np.array_equal(orc.median(img, r), orc.ref_median(img, r))


def test_quirks(orc):
    """Appendix A.2: threshold, -256 -> 0, search cut-off x+d > W, zero AD for x < d, lowest-d ties."""
    h, w, r, D = 12, 20, 2, 8
    L = np.full((h, w), 200, np.uint8)
    R = np.zeros((h, w), np.uint8)  # AD = 200 everywhere it is defined -> full windows fail the 50*N threshold
    d = orc.sad_wta(L, R, r, D)
    # interior pixels far from the x<d zero region: nothing accepted -> 0
    assert d[6, 15] == 0
    # identical images: every d has SAD 0 only at d = 0 ... ties resolve to the lowest d
    L = np.arange(h * w, dtype=np.uint8).reshape(h, w)
    assert np.all(orc.sad_wta(L, L, r, D)[:, r + D:] == 0)
    # AD slice zero region
    rng = np.random.default_rng(1)
    L = rng.integers(0, 256, (h, w), dtype=np.uint8); R = rng.integers(0, 256, (h, w), dtype=np.uint8)
    vol = orc.ad_volume(L, R, D)
    for dd in range(D):
        assert np.all(vol[dd, :, :dd] == 0)
        assert np.array_equal(vol[dd, :, dd:], np.abs(L[:, dd:].astype(int) - R[:, :w - dd].astype(int)).astype(np.uint8))


def _gf_naive(I, p, r, eps):
    """Pure-numpy float64 GF-v1 with explicit clipped windows (tiny inputs only)."""
    h, w = I.shape
    I = I.astype(np.float64); p = p.astype(np.float64)
    a = np.zeros((h, w)); b = np.zeros((h, w)); N = np.zeros((h, w))
    for y in range(h):
        for x in range(w):
            ys = slice(max(0, y - r), min(h, y + r + 1)); xs = slice(max(0, x - r), min(w, x + r + 1))
            wi, wp = I[ys, xs], p[ys, xs]
            n = wi.size
            mi, mp = wi.mean(), wp.mean()
            var = (wi * wi).mean() - mi * mi
            cov = (wi * wp).mean() - mi * mp
            a[y, x] = cov / (var + eps)
            b[y, x] = mp - a[y, x] * mi
            N[y, x] = n
    q = np.zeros((h, w))
    for y in range(h):
        for x in range(w):
            ys = slice(max(0, y - r), min(h, y + r + 1)); xs = slice(max(0, x - r), min(w, x + r + 1))
            q[y, x] = (a[ys, xs].sum() * I[y, x] + b[ys, xs].sum()) / N[y, x]
    return q


def test_gf_matches_naive_definition(orc):
    rng = np.random.default_rng(3)
    h, w, r, D = 14, 19, 2, 6
    L = rng.integers(0, 256, (h, w), dtype=np.uint8); R = rng.integers(0, 256, (h, w), dtype=np.uint8)
    for view in (0, 1):
        q = orc.gf_cost_slices(L, R, r, 0, D, eps=6.5025, view=view)
        for d in range(D):
            p = orc.ad_slice(L, R, d, view)
            ref = _gf_naive(L if view == 0 else R, p, r, 6.5025)
            assert np.allclose(q[d], ref, rtol=1e-9, atol=1e-9)
    disp, cost = orc.gf_wta(L, R, r, D, return_cost=True)
    q = orc.gf_cost_slices(L, R, r, 0, D)
    assert np.array_equal(disp, np.argmin(q, axis=0).astype(np.uint8))  # first minimum wins
    assert np.allclose(cost, q.min(axis=0))


def test_right_view_and_lr(orc):
    """StereoHelper.cpp:156-180 (right volume) and StereoDisparity.cpp:136-147 (LR check), literal loops."""
    rng = np.random.default_rng(5)
    h, w, D = 6, 13, 9
    L = rng.integers(0, 256, (h, w), dtype=np.uint8); R = rng.integers(0, 256, (h, w), dtype=np.uint8)
    left = np.stack([orc.ad_slice(L, R, d, 0) for d in range(D)], axis=-1).astype(int)  # [y][x][d]
    right = left.copy()
    for y in range(h):
        for x in range(w):
            for d in range(D):
                if x + d < w:
                    right[y, x, d] = np.abs(int(L[y, x + d]) - int(R[y, x]))  # == leftPtr(y, x+d, d): x+d >= d
                else:
                    right[y, x, d] = right[y, x, d - 1]
    for d in range(D):
        assert np.array_equal(orc.ad_slice(L, R, d, 1), right[:, :, d].astype(np.uint8))
    DL = rng.integers(0, 6, (h, w), dtype=np.uint8); DR = rng.integers(0, 6, (h, w), dtype=np.uint8)
    occ, mask = orc.lr_check(DL, DR)
    for y in range(h):
        for x in range(w):
            d = int(DL[y, x])
            e = 1 if x - d < 0 else int(d == 0 or abs(d - int(DR[y, x - d])) > 1)
            assert occ[y, x] == e and mask[y, x] == (not e)


def test_median_is_replicate_border(orc):
    from scipy.ndimage import median_filter
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, (33, 47), dtype=np.uint8)
    for r in (1, 2, 3):
        assert np.array_equal(orc.median(img, r), median_filter(img, size=2 * r + 1, mode="nearest"))


def test_packed_min_equivalence(fx, orc):
    """Appendix A.2: any partition of the d range combined by min over (SAD<<8|d), starting from the init
    word (50*(2r+1)^2)<<8, reproduces getDisp -- the identity the CTAs and GPUs rely on."""
    L, R = fx["ArtDemo_L"][:96, :160].copy(), fx["ArtDemo_R"][:96, :160].copy()
    r, D = 5, 64
    h, w = L.shape
    init = (50 * (2 * r + 1) ** 2) << 8
    xs = np.arange(w)[None, :]
    parts = []
    for k in range(4):  # 4-way interleaved split
        keys = np.full((h, w), init, np.int64)
        for d in range(k, D, 4):
            sad = orc.sad_slice(L, R, r, d).astype(np.int64)
            cand = np.where(xs + d <= w, (sad << 8) | d, np.int64(2 ** 62))
            keys = np.minimum(keys, cand)
        parts.append(keys)
    disp = (np.minimum.reduce(parts) & 0xFF).astype(np.uint8)
    assert np.array_equal(disp, orc.sad_wta(L, R, r, D))


def test_remap_and_cvtcolor_restatements(orc):
    """CPU_Remap (Utility.cpp:236-264) and kernalCvtColor / cvtColor_cpu, vs float32 numpy."""
    rng = np.random.default_rng(13)
    h, w = 37, 53
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    mx = (rng.random((h, w), dtype=np.float32) * (w + 6) - 3).astype(np.float32)
    my = (rng.random((h, w), dtype=np.float32) * (h + 6) - 3).astype(np.float32)
    out = orc.remap(img, mx, my)
    f = np.float32
    x, y = my, mx  # the interpolator's x is the row coordinate
    x1 = np.floor(x).astype(np.int64); y1 = np.floor(y).astype(np.int64)
    ok = ~((x1 < 0) | (x1 + 1 >= h) | (y1 < 0) | (y1 + 1 >= w))
    xc, yc = np.clip(x1, 0, h - 2), np.clip(y1, 0, w - 2)
    Q11, Q12 = img[xc, yc].astype(f), img[xc, yc + 1].astype(f)
    Q21, Q22 = img[xc + 1, yc].astype(f), img[xc + 1, yc + 1].astype(f)
    wx2 = ((x1 + 1).astype(f) - x).astype(f); wx1 = (x - x1.astype(f)).astype(f)
    left = ((wx2 * Q11).astype(f) + (wx1 * Q21).astype(f)).astype(f)
    right = ((wx2 * Q12).astype(f) + (wx1 * Q22).astype(f)).astype(f)
    res = ((((y1 + 1).astype(f) - y).astype(f) * left).astype(f) + ((y - y1.astype(f)).astype(f) * right).astype(f)).astype(f)
    ref = np.where(ok, np.clip(np.rint(res), 0, 255), 0).astype(np.uint8)  # np.rint rounds half to even
    assert np.array_equal(out, ref)
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    s = ((f(.299) * rgb[..., 0].astype(f)).astype(f) + (f(.587) * rgb[..., 1].astype(f)).astype(f)).astype(f)
    s = (s + (f(.114) * rgb[..., 2].astype(f)).astype(f)).astype(f)
    assert np.array_equal(orc.cvtcolor(rgb, truncate=True), s.astype(np.uint8))
    # the GPU kernel's expression is contracted by nvcc: fma(.114, c2, fma(.299, c0, .587 * c1)); float64 holds the
    # products and sums of these operands exactly, so one rounding to float32 emulates each fma
    d = np.float64
    t = (f(.587) * rgb[..., 1].astype(f)).astype(f)
    t = (d(f(.299)) * rgb[..., 0].astype(d) + t.astype(d)).astype(f)
    t = (d(f(.114)) * rgb[..., 2].astype(d) + t.astype(d)).astype(f)
    assert np.array_equal(orc.cvtcolor(rgb, truncate=False), np.clip(np.rint(t), 0, 255).astype(np.uint8))


def test_right_view_and_float_wta_against_compiled_stmatching(orc):
    """SURVEY 8a rows a6 / a7 pinned on the reference's own code: STMatching/StereoHelper.cpp compiled unmodified
    (oracle/_ref/libstref.so).  (i) GetRightMatchingCostFromLeft on the left AD volume == the oracle's right-view AD
    slices; (ii) GetDisparity_WTA == "strict <, first minimum wins" on volumes with exact ties, which is the rule
    orc_gf_wta applies to its own costs."""
    if not orc.have_stref():
        pytest.skip("oracle/_ref/libstref.so not built (reference tree absent)")
    rng = np.random.default_rng(11)
    # incl. D == w (every column is an edge column); the reference indexes out of bounds for D > w (x = w - D < 0)
    for (h, w, D) in ((7, 19, 9), (5, 40, 33), (9, 12, 12), (4, 16, 15)):
        L = rng.integers(0, 256, (h, w), dtype=np.uint8); R = rng.integers(0, 256, (h, w), dtype=np.uint8)
        left = np.stack([orc.ad_slice(L, R, d, 0) for d in range(D)], axis=-1).astype(np.float32)   # [y][x][d]
        right = orc.ref_right_from_left(left)
        ours = np.stack([orc.ad_slice(L, R, d, 1) for d in range(D)], axis=-1).astype(np.float32)
        assert np.array_equal(right, ours), (h, w, D)
        vol = rng.integers(0, 4, (h, w, D)).astype(np.float32) * 0.25 - 0.25     # many exact ties, negative values
        assert np.array_equal(orc.ref_wta_float(vol), np.argmin(vol, axis=-1).astype(np.uint8))
    # the same rule on real guided-filter costs: the oracle's WTA picks what the reference's WTA picks
    Lr = rng.integers(0, 256, (24, 40), dtype=np.uint8); Rr = np.roll(Lr, -3, axis=1)
    q = orc.gf_cost_slices(Lr, Rr, 3, 0, 12)                      # float64 [D][H][W]
    d_or = orc.gf_wta(Lr, Rr, 3, 12)
    d_ref = orc.ref_wta_float(np.ascontiguousarray(np.moveaxis(q, 0, -1)).astype(np.float32))
    diff = d_or != d_ref                                           # may differ only where the float32 cast creates a tie
    q32 = np.moveaxis(q, 0, -1).astype(np.float32)
    ys, xs = np.nonzero(diff)
    assert all(q32[y, x, d_or[y, x]] == q32[y, x, d_ref[y, x]] for y, x in zip(ys, xs))


def test_remap_and_cvtcolor_against_compiled_utility(orc):
    """SURVEY 8f rows 1 / 2 pinned on the reference's own code: BlockMatching/Utility.cpp compiled unmodified
    (oracle/_ref/libutilref.so).  CPU_Remap (incl. its (ycoo, xcoo) argument order, the out-of-image rule and the
    rounding of saturate_cast) and cvtColor_cpu (truncation) == the oracle's restatements, bit for bit."""
    if not orc.have_utilref():
        pytest.skip("oracle/_ref/libutilref.so not built (reference tree absent)")
    rng = np.random.default_rng(17)
    for (h, w) in ((37, 53), (200, 320), (5, 4)):
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        mx = (rng.random((h, w), dtype=np.float32) * (w + 6) - 3).astype(np.float32)
        my = (rng.random((h, w), dtype=np.float32) * (h + 6) - 3).astype(np.float32)
        # exact integers and .5 offsets exercise the borders (x2 >= rows -> 0) and the ties of the rounding
        mx.flat[::7] = np.round(mx.flat[::7]); my.flat[::5] = np.round(my.flat[::5]) + 0.5
        assert np.array_equal(orc.ref_cpu_remap(img, mx, my), orc.remap(img, mx, my)), (h, w)
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(orc.ref_cvtcolor_cpu(rgb), orc.cvtcolor(rgb, truncate=True)), (h, w)
    # identity maps reproduce the image except the last row / column (x2 >= rows, y2 >= cols -> 0): a reference quirk
    img = rng.integers(0, 256, (9, 11), dtype=np.uint8)
    yy, xx = np.mgrid[0:9, 0:11].astype(np.float32)
    out = orc.ref_cpu_remap(img, xx, yy)
    assert np.array_equal(out[:-1, :-1], img[:-1, :-1]) and not out[-1].any() and not out[:, -1].any()


def _art_demo_bgr():
    import cv2
    d = os.path.join(ROOT, "tests", "golden", "art_demo")
    return cv2.imread(os.path.join(d, "view1_.png")), cv2.imread(os.path.join(d, "view5_.png"))


def test_segment_tree_host_builder_equals_reference_tree():
    """SURVEY 8f row 4, host stage: the product's O(N) tree builder (gsm_st_build_tree_host: counting sort + Kruskal with
    the adaptive threshold + breadth-first ordering) reproduces the ordered tree of the reference's BuildSegmentTree
    (SegmentTree.cpp:38-139, compiled unmodified into oracle/_ref/libsegref.so) node for node -- same order, same
    fathers, same quantised edge weights -- on the reference's demo image, on synthetic colour images, on odd shapes,
    on a constant image (all weights tie) and for several tau.  No GPU involved."""
    from oracle import oracle as O
    if not O.have_segref():
        pytest.skip("oracle/_ref/libsegref.so not built (needs the reference tree)")
    import gpu_stereo_matching_b200 as g
    from gpu_stereo_matching_b200 import data as gdata
    L, _ = _art_demo_bgr()
    cases = [(L, 1200.0), (L, 300.0), (L[:57, :83].copy(), 1200.0), (gdata.synthetic_color_pair(120, 200, 5)[0], 1200.0),
             (np.full((40, 30, 3), 77, np.uint8), 1200.0), (L[:3, :50].copy(), 1200.0), (L[:50, :3].copy(), 50.0)]  # the reference's ctmf needs >= 3x3
    rng = np.random.default_rng(0)
    cases.append((rng.integers(0, 256, (64, 96, 3), dtype=np.uint8), 1200.0))  # noise: heavy use of the second pass
    for img, tau in cases:
        _, order, father, fdist = O.ref_st_filter(img, None, 0.1, tau)
        wr, wu = O.st_edge_weights(img)
        o2, f2, d2, levels = g.st_build_tree_host(wr, wu, tau)
        assert np.array_equal(order, o2) and np.array_equal(father, f2) and np.array_equal(fdist, d2), (img.shape, tau)
        assert levels >= 1
