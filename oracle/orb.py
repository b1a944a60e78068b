"""CPU restatement of cv2.ORB_create(nfeatures).detectAndCompute(gray, None) -- TEST INFRASTRUCTURE, not product code.

The reference detects ORB features with `cv2.ORB_create(max_features)` / `detector.detectAndCompute(scaled, None)`
(scripts/lib/image.py:243-245, :324).  The arithmetic lives in OpenCV (features2d/orb.cpp, fast.cpp; pinned
`opencv = 4.0.1` in environment.yml, 4.13.0 in this image), which is not vendored under /root/reference, so this
file restates the published ORB pipeline (Rublee et al. 2011; OpenCV's defaults: 8 levels, scale 1.2, edge
threshold 31, patch 31, FAST threshold 20, Harris score, WTA_K 2) stage by stage.  Every stage was pinned against
live cv2 in the build container (tests/test_orb.py, tests/golden/make_golden_orb.py):

  pyramid      cv2.resize(prev_level, size, INTER_LINEAR_EXACT), each level from the previous one -- bit exact
               (8.8 fixed-point weights per axis, horizontal then vertical, round at 2^15)
  FAST-9/16    cv2.FastFeatureDetector_create(20, True): positions and scores bit exact
  retainBest   keep everything whose response is >= the n-th largest (ties survive)
  Harris       7x7 block Sobel sums, float32 response, same evaluation order as OpenCV
  orientation  intensity centroid over the circular patch (radius 15) + OpenCV's polynomial fastAtan2
  descriptor   steered BRIEF on the level blurred by the separable FLOAT 7-tap Gaussian (sigma 2) that ORB's in-place
               GaussianBlur call resolves to (NOT the fixed-point path a stand-alone 8-bit GaussianBlur takes), the
               256 test pairs of OpenCV's learned pattern (recovered by probing cv2.ORB.compute with step images:
               tests/golden/make_golden_orb.py documents the procedure; the table is tests/golden/orb_pattern.npy)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import math
import os

import numpy as np

# ORB_create's scaleFactor parameter is a float (1.2f) stored in a double member: 1.2000000476837158
NLEVELS, SCALE_FACTOR, EDGE_THRESHOLD, PATCH_SIZE, FAST_THRESHOLD, HARRIS_K = 8, float(np.float32(1.2)), 31, 31, 20, 0.04
HALF_PATCH = PATCH_SIZE // 2
CIRCLE = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0),
          (-3, 1), (-2, 2), (-1, 3)]
_HERE = os.path.dirname(os.path.abspath(__file__))


def pattern() -> np.ndarray:
    """[256, 4] int: (x0, y0, x1, y1) of every binary test (OpenCV's bit_pattern_31_, recovered by probing)."""
    return np.load(os.path.join(_HERE, "..", "tests", "golden", "orb_pattern.npy")).astype(np.int64)


def gaussian_kernel7() -> np.ndarray:
    """cv2.getGaussianKernel(7, 2, CV_32F): exp(-x^2 / (2 sigma^2)) normalised in double, stored as float32."""
    x = np.arange(-3, 4, dtype=np.float64)
    k = np.exp(-(x * x) / 8.0)
    return (k / k.sum()).astype(np.float32)


def cv_round(x) -> int:
    return int(np.rint(x))          # round half to even, like cvRound


def level_scales(nlevels=NLEVELS, factor=SCALE_FACTOR):
    return [np.float32(math.pow(factor, l)) for l in range(nlevels)]


def features_per_level(nfeatures, nlevels=NLEVELS, factor=SCALE_FACTOR):
    """orb.cpp computeKeyPoints: geometric split of nfeatures over the levels (float32 arithmetic)."""
    f = np.float32(1.0 / factor)
    nd = np.float32(nfeatures) * (np.float32(1) - f) / (np.float32(1) - np.float32(math.pow(float(f), float(nlevels))))
    out, total = [], 0
    for _ in range(nlevels - 1):
        out.append(cv_round(nd))
        total += out[-1]
        nd = np.float32(nd * f)
    out.append(max(nfeatures - total, 0))
    return out


def resize_linear_exact(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT) for uint8, bit exact."""
    sh, sw = src.shape

    def coef(dn, sn):
        scale = sn / dn
        f = (np.arange(dn) + 0.5) * scale - 0.5
        i = np.floor(f).astype(np.int64)
        a = f - i
        lo, hi = i < 0, i >= sn - 1
        i = np.where(lo, 0, np.where(hi, sn - 1, i))
        a = np.where(lo | hi, 0.0, a)
        return i, np.rint(a * 256).astype(np.int64)
    xi, xa = coef(dw, sw)
    yi, ya = coef(dh, sh)
    s = src.astype(np.int64)
    h = s[:, xi] * (256 - xa) + s[:, np.minimum(xi + 1, sw - 1)] * xa
    v = h[yi] * (256 - ya)[:, None] + h[np.minimum(yi + 1, sh - 1)] * ya[:, None]
    return ((v + (1 << 15)) >> 16).astype(np.uint8)


def build_pyramid(gray: np.ndarray, nlevels=NLEVELS):
    scales = level_scales(nlevels)
    levels = [gray]
    for l in range(1, nlevels):
        w = cv_round(np.float32(gray.shape[1]) / scales[l])
        h = cv_round(np.float32(gray.shape[0]) / scales[l])
        levels.append(resize_linear_exact(levels[-1], w, h))
    return levels, scales


def fast_scores(img: np.ndarray, t=FAST_THRESHOLD) -> np.ndarray:
    """FAST-9/16 corner score of every pixel (0 = not a corner), fast.cpp cornerScore<16>."""
    h, w = img.shape
    out = np.zeros((h, w), np.int32)
    if h < 7 or w < 7:
        return out
    v = img[3:h - 3, 3:w - 3].astype(np.int32)
    d = np.stack([v - img[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx].astype(np.int32) for dx, dy in CIRCLE])
    d2 = np.concatenate([d, d[:9]])
    A = np.full(v.shape, -10 ** 6)
    B = np.full(v.shape, -10 ** 6)
    for s in range(16):
        arc = d2[s:s + 9]
        A = np.maximum(A, arc.min(0))
        B = np.maximum(B, (-arc).min(0))
    m = np.maximum(A, B)
    out[3:h - 3, 3:w - 3] = np.where(m > t, m - 1, 0)
    return out


def fast_detect(img: np.ndarray, t=FAST_THRESHOLD):
    """cv2.FastFeatureDetector_create(t, True).detect: (x, y, score) with 3x3 non-maximum suppression."""
    s = fast_scores(img, t)
    h, w = s.shape
    c = s[1:-1, 1:-1]
    keep = c > 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx or dy:
                keep &= c > s[1 + dy:h - 1 + dy, 1 + dx:w - 1 + dx]
    ys, xs = np.nonzero(keep)
    return xs + 1, ys + 1, c[ys, xs].astype(np.float32)


def retain_best(resp: np.ndarray, n: int) -> np.ndarray:
    """KeyPointsFilter::retainBest as a mask: everything >= the n-th largest response survives."""
    if n <= 0:
        return np.zeros(len(resp), bool)
    if len(resp) <= n:
        return np.ones(len(resp), bool)
    kth = np.partition(resp, len(resp) - n)[len(resp) - n]
    return resp >= kth


def reflect101(img: np.ndarray, b: int) -> np.ndarray:
    return np.pad(img, b, mode="reflect")


def harris_responses(padded: np.ndarray, b: int, xs, ys, block=7, k=HARRIS_K) -> np.ndarray:
    """orb.cpp HarrisResponses on the level with its reflect-101 border of width b."""
    r = block // 2
    img = padded.astype(np.int32)
    scale = np.float32(1.0) / (np.float32(4 * block) * np.float32(255.0))
    scale_sq_sq = np.float32(np.float32(scale * scale) * scale) * scale
    out = np.zeros(len(xs), np.float32)
    ii, jj = np.mgrid[-r:r + 1, -r:r + 1]
    for n, (x0, y0) in enumerate(zip(xs, ys)):
        yy, xx = y0 + b + ii, x0 + b + jj
        Ix = (img[yy, xx + 1] - img[yy, xx - 1]) * 2 + (img[yy - 1, xx + 1] - img[yy - 1, xx - 1]) + (img[yy + 1, xx + 1] - img[yy + 1, xx - 1])
        Iy = (img[yy + 1, xx] - img[yy - 1, xx]) * 2 + (img[yy + 1, xx - 1] - img[yy - 1, xx - 1]) + (img[yy + 1, xx + 1] - img[yy - 1, xx + 1])
        a, bb, c = np.float32(int((Ix * Ix).sum())), np.float32(int((Iy * Iy).sum())), np.float32(int((Ix * Iy).sum()))
        out[n] = (a * bb - c * c - np.float32(k) * (a + bb) * (a + bb)) * scale_sq_sq
    return out


def umax_table(half=HALF_PATCH):
    um = [0] * (half + 2)
    vmax = int(math.floor(half * math.sqrt(2.0) / 2 + 1))
    vmin = int(math.ceil(half * math.sqrt(2.0) / 2))
    for v in range(vmax + 1):
        um[v] = cv_round(math.sqrt(float(half) * half - v * v))
    v0 = 0
    for v in range(half, vmin - 1, -1):
        while um[v0] == um[v0 + 1]:
            v0 += 1
        um[v] = v0
        v0 += 1
    return um


_P1 = np.float32(0.9997878412794807 * (180 / math.pi))
_P3 = np.float32(-0.3258083974640975 * (180 / math.pi))
_P5 = np.float32(0.1555786518463281 * (180 / math.pi))
_P7 = np.float32(-0.04432655554792128 * (180 / math.pi))
_EPS = np.float32(2.220446049250313e-16)


def fast_atan2(y, x) -> np.float32:
    """OpenCV's fastAtan2 (degrees, float32 polynomial)."""
    y, x = np.float32(y), np.float32(x)
    ax, ay = np.float32(abs(x)), np.float32(abs(y))
    if ax >= ay:
        c = ay / (ax + _EPS)
        c2 = c * c
        a = (((_P7 * c2 + _P5) * c2 + _P3) * c2 + _P1) * c
    else:
        c = ax / (ay + _EPS)
        c2 = c * c
        a = np.float32(90.0) - (((_P7 * c2 + _P5) * c2 + _P3) * c2 + _P1) * c
    if x < 0:
        a = np.float32(180.0) - a
    if y < 0:
        a = np.float32(360.0) - a
    return np.float32(a)


def ic_angles(padded: np.ndarray, b: int, xs, ys, half=HALF_PATCH) -> np.ndarray:
    um = umax_table(half)
    img = padded.astype(np.int64)
    out = np.zeros(len(xs), np.float32)
    for n, (x0, y0) in enumerate(zip(xs, ys)):
        cy, cx = y0 + b, x0 + b
        u = np.arange(-half, half + 1)
        m10 = int((u * img[cy, cx - half:cx + half + 1]).sum())
        m01 = 0
        for v in range(1, half + 1):
            d = um[v]
            uu = np.arange(-d, d + 1)
            plus, minus = img[cy + v, cx - d:cx + d + 1], img[cy - v, cx - d:cx + d + 1]
            m01 += v * int((plus - minus).sum())
            m10 += int((uu * (plus + minus)).sum())
        out[n] = fast_atan2(np.float32(m01), np.float32(m10))
    return out


def blur_level(padded: np.ndarray) -> np.ndarray:
    """The blur ORB applies before the descriptors: separable float32 7-tap Gaussian, rows then columns, rounded to
    uint8 (the padded layer already carries its reflect-101 border; the outermost 3 pixels are not used)."""
    k = gaussian_kernel7()
    f = padded.astype(np.float32)
    h = np.zeros_like(f)
    for t in range(7):
        h[:, 3:-3] += k[t] * f[:, t:f.shape[1] - 6 + t]
    v = np.zeros_like(f)
    for t in range(7):
        v[3:-3] += k[t] * h[t:f.shape[0] - 6 + t]
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def descriptors(blurred: np.ndarray, b: int, xs, ys, angles_deg) -> np.ndarray:
    P = pattern()
    out = np.zeros((len(xs), 32), np.uint8)
    for n, (x0, y0, ang) in enumerate(zip(xs, ys, angles_deg)):
        ang = np.float32(ang) * np.float32(math.pi / 180.0)
        a, s = np.float32(math.cos(float(ang))), np.float32(math.sin(float(ang)))
        px0, py0, px1, py1 = (P[:, i].astype(np.float32) for i in range(4))
        ix0 = np.rint(px0 * a - py0 * s).astype(np.int64)
        iy0 = np.rint(px0 * s + py0 * a).astype(np.int64)
        ix1 = np.rint(px1 * a - py1 * s).astype(np.int64)
        iy1 = np.rint(px1 * s + py1 * a).astype(np.int64)
        t0 = blurred[y0 + b + iy0, x0 + b + ix0]
        t1 = blurred[y0 + b + iy1, x0 + b + ix1]
        out[n] = np.packbits((t0 < t1).astype(np.uint8), bitorder="little")
    return out


def detect_and_compute(gray: np.ndarray, nfeatures: int = 500):
    """cv2.ORB_create(nfeatures).detectAndCompute(gray, None) -> dict of arrays
    (pt [n,2] f32, size, angle, response, octave, des [n,32] u8).  Key points are listed level by level; inside a
    level in raster order (OpenCV's order inside a level comes out of std::nth_element and is not specified)."""
    levels, scales = build_pyramid(gray)
    per_level = features_per_level(nfeatures)
    border = max(EDGE_THRESHOLD, int(math.ceil(HALF_PATCH * math.sqrt(2.0))), 9 // 2) + 1
    res = dict(pt=[], size=[], angle=[], response=[], octave=[], des=[])
    for l, img in enumerate(levels):
        h, w = img.shape
        xs, ys, sc = fast_detect(img)
        inside = (xs >= EDGE_THRESHOLD) & (xs < w - EDGE_THRESHOLD) & (ys >= EDGE_THRESHOLD) & (ys < h - EDGE_THRESHOLD)
        xs, ys, sc = xs[inside], ys[inside], sc[inside]
        keep = retain_best(sc, 2 * per_level[l])
        xs, ys = xs[keep], ys[keep]
        padded = reflect101(img, border)
        hr = harris_responses(padded, border, xs, ys)
        keep = retain_best(hr, per_level[l])
        xs, ys, hr = xs[keep], ys[keep], hr[keep]
        ang = ic_angles(padded, border, xs, ys)
        des = descriptors(blur_level(padded), border, xs, ys, ang)
        sf = scales[l]
        res["pt"].append(np.stack([xs.astype(np.float32) * sf, ys.astype(np.float32) * sf], 1).astype(np.float32))
        res["size"].append(np.full(len(xs), np.float32(PATCH_SIZE) * sf, np.float32))
        res["angle"].append(ang)
        res["response"].append(hr)
        res["octave"].append(np.full(len(xs), l, np.int32))
        res["des"].append(des)
    return {k: np.concatenate(v) for k, v in res.items()}
