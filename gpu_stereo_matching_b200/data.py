"""Seeded synthetic stereo pairs of the BASELINE shapes -- data only, no compute path."""
from __future__ import annotations

import numpy as np


def _smooth_texture(rng, h, w):
    """Sum of 3 octaves of bilinearly up-sampled uniform noise, scaled to 0..255."""
    img = np.zeros((h, w), np.float32)
    for octave, amp in ((8, 0.5), (32, 0.3), (128, 0.2)):
        gh, gw = h // octave + 2, w // octave + 2
        g = rng.random((gh, gw), dtype=np.float32)
        ys = np.arange(h, dtype=np.float32) / octave
        xs = np.arange(w, dtype=np.float32) / octave
        y0, x0 = ys.astype(np.int32), xs.astype(np.int32)
        fy, fx = (ys - y0)[:, None], (xs - x0)[None, :]
        a = g[y0][:, x0]; b = g[y0][:, x0 + 1]; c = g[y0 + 1][:, x0]; d = g[y0 + 1][:, x0 + 1]
        img += amp * ((1 - fy) * ((1 - fx) * a + fx * b) + fy * ((1 - fx) * c + fx * d))
    return img * 255.0


def synthetic_pair(h: int, w: int, seed: int, dmin: int = 4, dmax: int = 120, noise: float = 2.0):
    """Left = smooth texture, right = left warped by a piecewise-planar disparity field in [dmin, dmax]
    plus N(0, noise) -- returns (left u8, right u8, true disparity f32)."""
    rng = np.random.default_rng(seed)
    left = _smooth_texture(rng, h, w)
    disp = np.full((h, w), float(dmin), np.float32)
    for _ in range(6):  # a few fronto-parallel / slanted planes
        y0, x0 = rng.integers(0, h), rng.integers(0, w)
        hh, ww = rng.integers(h // 6, h // 2), rng.integers(w // 6, w // 2)
        base = rng.uniform(dmin, dmax)
        slope = rng.uniform(-0.02, 0.02)
        ys = slice(y0, min(h, y0 + hh)); xs = slice(x0, min(w, x0 + ww))
        plane = base + slope * (np.arange(xs.start, xs.stop, dtype=np.float32) - x0)[None, :]
        disp[ys, xs] = np.clip(plane, dmin, dmax)
    xr = np.arange(w, dtype=np.float32)[None, :] + disp  # right(x) = left(x + d)
    x0 = np.clip(np.floor(xr).astype(np.int32), 0, w - 1)
    x1 = np.clip(x0 + 1, 0, w - 1)
    f = xr - np.floor(xr)
    rows = np.arange(h)[:, None]
    right = (1 - f) * left[rows, x0] + f * left[rows, x1]
    right += rng.normal(0.0, noise, (h, w)).astype(np.float32)
    to_u8 = lambda a: np.clip(np.rint(a), 0, 255).astype(np.uint8)
    return to_u8(left), to_u8(right), disp


def synthetic_color_pair(h: int, w: int, seed: int, dmax: int = 40):
    """3-channel (B, G, R) version of synthetic_pair: three textures sharing ONE disparity field -> (left, right) u8
    [h, w, 3].  Input of the segment-tree stereo."""
    rng = np.random.default_rng(seed)
    L0, R0, disp = synthetic_pair(h, w, seed, dmax=dmax)
    left = np.empty((h, w, 3), np.uint8)
    right = np.empty((h, w, 3), np.uint8)
    left[:, :, 0], right[:, :, 0] = L0, R0
    xr = np.arange(w, dtype=np.float32)[None, :] + disp
    x0 = np.clip(np.floor(xr).astype(np.int32), 0, w - 1)
    x1 = np.clip(x0 + 1, 0, w - 1)
    f = xr - np.floor(xr)
    rows = np.arange(h)[:, None]
    for c in (1, 2):
        tex = _smooth_texture(rng, h, w)
        left[:, :, c] = np.clip(np.rint(tex), 0, 255).astype(np.uint8)
        r = (1 - f) * tex[rows, x0] + f * tex[rows, x1] + rng.normal(0.0, 2.0, (h, w)).astype(np.float32)
        right[:, :, c] = np.clip(np.rint(r), 0, 255).astype(np.uint8)
    return left, right


def synthetic_batch(n: int, h: int, w: int, seed0: int, **kw):
    L = np.empty((n, h, w), np.uint8); R = np.empty((n, h, w), np.uint8)
    for i in range(n):
        L[i], R[i], _ = synthetic_pair(h, w, seed0 + i, **kw)
    return L, R


def noise_pair(h: int, w: int, seed: int):
    """White-noise pair: worst case / tie-stress parity input."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (h, w), dtype=np.uint8), rng.integers(0, 256, (h, w), dtype=np.uint8)


# Calibration constants of the reference rig: the VALUES of /root/reference/Calib_Data_OpenCV.yml
# (keys read by LoadDataBatch, BlockMatching/Utility.cpp:29-40).  Data, not code; kept here because the
# reference tree does not exist on the GPU box.  "RotationVec" is a 3x3 rotation matrix despite its name.
CALIB = {
    "LeftMat": [[1116.744104, 0.0, 624.050472], [0.0, 1114.049167, 372.767559], [0.0, 0.0, 1.0]],
    "RightMat": [[1125.408483, 0.0, 631.532381], [0.0, 1122.554172, 363.269831], [0.0, 0.0, 1.0]],
    "LeftDist": [0.036791, -0.216298, -0.002058, -0.000422, 0.0],
    "RightDist": [0.047295, -0.311851, 0.000326, -0.001401, 0.0],
    "RotationVec": [[0.999831, 0.004900, -0.017693], [-0.004914, 0.999988, -0.000744],
                    [0.017690, 0.000831, 0.999843]],
    "TranslationVec": [-46.993557, -0.107737, -0.240733],
}


def rectify_maps(width: int = 1280, height: int = 720):
    """Host-side rectification maps exactly as Rectify() builds them (BlockMatching/Utility.cpp:228-234):
    stereoRectify(..., CALIB_ZERO_DISPARITY) + 2 x initUndistortRectifyMap(CV_32FC1).  Stays on the host
    (north_star); returns (mapX1, mapY1, mapX2, mapY2)."""
    import cv2
    K1 = np.array(CALIB["LeftMat"], np.float64); K2 = np.array(CALIB["RightMat"], np.float64)
    D1 = np.array(CALIB["LeftDist"], np.float64); D2 = np.array(CALIB["RightDist"], np.float64)
    Rm = np.array(CALIB["RotationVec"], np.float64); T = np.array(CALIB["TranslationVec"], np.float64)
    size = (width, height)
    R1, R2, P1, P2, Q, _, _ = cv2.stereoRectify(K1, D1, K2, D2, size, Rm, T, flags=cv2.CALIB_ZERO_DISPARITY)
    m1x, m1y = cv2.initUndistortRectifyMap(K1, D1, R1, P1, size, cv2.CV_32FC1)
    m2x, m2y = cv2.initUndistortRectifyMap(K2, D2, R2, P2, size, cv2.CV_32FC1)
    return m1x, m1y, m2x, m2y


def rectified_stream(n: int, seed0: int = 1234, width: int = 1280, height: int = 720, rectify: bool = True):
    """BASELINE config 3 input: n synthetic 1280x720 pairs (seeds seed0+f) pushed through the rig's
    rectification maps on the host (cv2.remap INTER_LINEAR).  Falls back to un-rectified frames when cv2 is
    unavailable (noted by the second return value)."""
    L, R = synthetic_batch(n, height, width, seed0)
    if not rectify:
        return L, R, False
    try:
        import cv2
        m1x, m1y, m2x, m2y = rectify_maps(width, height)
    except Exception:
        return L, R, False
    for i in range(n):
        L[i] = cv2.remap(L[i], m1x, m1y, cv2.INTER_LINEAR)
        R[i] = cv2.remap(R[i], m2x, m2y, cv2.INTER_LINEAR)
    return L, R, True
