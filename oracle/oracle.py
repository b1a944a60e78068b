"""oracle.py — CPU restatement of the reference's pairwise matching path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module, and
only as the checker or as the timed CPU baseline.  Nothing under
imageanalysis_b200/ imports it; the product fails loudly when its CUDA
library is missing instead of falling back to this code.

Every function cites the reference lines it follows (paths relative to the
reference repo root).  The k-NN arithmetic itself lives in OpenCV, which the
reference does not vendor (environment.yml:13 pins opencv 4.0.1; this image
ships 4.13.0): `knn()` restates BFMatcher.knnMatch and is PINNED against
live cv2 output committed under tests/golden/ (tests/golden/make_golden.py
is the generator).  The Python-level reductions are restated from
scripts/lib/matcher.py and pinned against the reference module itself,
imported with shims by the same generator.
"""
from __future__ import annotations

import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor
from typing import List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "liboracle_knn.so")
_lib = None

NORM_L2, NORM_HAMMING = 0, 1


def _load():
    global _lib
    if _lib is None and os.path.exists(_SO):
        lib = C.CDLL(_SO)
        for name in ("oracle_knn_l2_u8", "oracle_knn_l2_f32", "oracle_knn_hamming"):
            getattr(lib, name).restype = None
            getattr(lib, name).argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                           C.c_void_p]
        _lib = lib
    return _lib


def have_c_oracle() -> bool:
    return _load() is not None


# --------------------------------------------------------------------------
# k-NN  (matcher.py:203-216 raw_matches -> cv2 BFMatcher.knnMatch)
# --------------------------------------------------------------------------
def knn_numpy(q: np.ndarray, t: np.ndarray, k: int, norm: int) -> Tuple[np.ndarray, np.ndarray]:
    """Pure-numpy restatement (small sizes).  Returns (idx [N,k] int32, dist [N,k] float32).
    Ordering: ascending distance, ties by ascending train index (stable argsort)."""
    if norm == NORM_L2:
        qa = q.astype(np.float64)
        ta = t.astype(np.float64)
        d2 = ((qa[:, None, :] - ta[None, :, :]) ** 2).sum(-1)
    else:
        x = np.bitwise_xor(q[:, None, :], t[None, :, :])
        d2 = np.unpackbits(x, axis=-1).sum(-1).astype(np.float64)
    order = np.argsort(d2, axis=1, kind="stable")[:, :k]
    dsel = np.take_along_axis(d2, order, axis=1)
    dist = (np.sqrt(dsel) if norm == NORM_L2 else dsel).astype(np.float32)
    idx = order.astype(np.int32)
    if idx.shape[1] < k:  # fewer train rows than k
        pad = k - idx.shape[1]
        idx = np.concatenate([idx, np.full((idx.shape[0], pad), -1, np.int32)], 1)
        dist = np.concatenate([dist, np.full((dist.shape[0], pad), np.inf, np.float32)], 1)
    return idx, dist


def knn(q: np.ndarray, t: np.ndarray, k: int, norm: int, threads: int = 1) -> Tuple[np.ndarray, np.ndarray]:
    """C restatement (oracle_knn.c) fanned out over `threads` row blocks; falls
    back to knn_numpy when the shared object has not been built."""
    lib = _load()
    if lib is None:
        return knn_numpy(q, t, k, norm)
    q = np.ascontiguousarray(q)
    t = np.ascontiguousarray(t)
    nq, dim = q.shape
    nt = t.shape[0]
    idx = np.empty((nq, k), np.int32)
    dist = np.empty((nq, k), np.float32)
    if norm == NORM_HAMMING:
        fn = lib.oracle_knn_hamming
        assert q.dtype == np.uint8 and t.dtype == np.uint8
    elif q.dtype == np.uint8:
        fn = lib.oracle_knn_l2_u8
    else:
        q = q.astype(np.float32, copy=False)
        t = t.astype(np.float32, copy=False)
        fn = lib.oracle_knn_l2_f32

    def run(lo, hi):
        if hi > lo:
            fn(q[lo:hi].ctypes.data, hi - lo, t.ctypes.data, nt, dim, k, idx[lo:hi].ctypes.data, dist[lo:hi].ctypes.data)

    threads = max(1, min(threads, nq))
    if threads == 1:
        run(0, nq)
    else:
        bounds = np.linspace(0, nq, threads + 1).astype(int)
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda b: run(int(b[0]), int(b[1])), zip(bounds[:-1], bounds[1:])))
    return idx, dist


# --------------------------------------------------------------------------
# reductions  (matcher.py:218-273, 187-200, 157-182)
# --------------------------------------------------------------------------
def reduce_ref_metric(idx: np.ndarray, dist: np.ndarray, match_ratio: float, max_distance: float, cap: int = 2000,
                      min_pairs: int = 25) -> List[List[int]]:
    """matcher.py:253-273.  metric = d0 * (d0/d1) in Python floats, stable
    ascending sort, keep metric < max_distance*match_ratio, clip to `cap`,
    gate on min_pairs.  Rows whose d1 is 0 would raise ZeroDivisionError in
    the reference; they are dropped here and on the device (documented)."""
    by_metric = []
    for qi in range(idx.shape[0]):
        if idx[qi, 0] < 0 or idx[qi, 1] < 0:
            continue
        d0 = float(dist[qi, 0])
        d1 = float(dist[qi, 1])
        if d1 == 0.0:
            continue
        ratio = d0 / d1
        metric = d0 * ratio
        by_metric.append([metric, qi, int(idx[qi, 0])])
    by_metric = sorted(by_metric, key=lambda f: f[0])
    out = [[qi, ti] for metric, qi, ti in by_metric if metric < max_distance * match_ratio]
    if len(out) > cap:
        out = out[:cap]
    if len(out) < min_pairs:
        return []
    return out


def reduce_lowe(idx: np.ndarray, dist: np.ndarray, match_ratio: float, cap: int = 2000,
                min_pairs: int = 0) -> List[List[int]]:
    """Plain Lowe gate, matcher.py:227 (`m[0].distance <= m[1].distance * match_ratio`),
    same test as find_obj.py:50-60 filter_matches."""
    out = []
    for qi in range(idx.shape[0]):
        if idx[qi, 0] < 0 or idx[qi, 1] < 0:
            continue
        if float(dist[qi, 0]) <= float(dist[qi, 1]) * match_ratio:
            out.append([qi, int(idx[qi, 0])])
    out = out[:cap]
    if len(out) < min_pairs:
        return []
    return out


def filter_cross_check(p1: Sequence[Sequence[int]], p2: Sequence[Sequence[int]]):
    """matcher.py:187-200."""
    s2 = {(int(a), int(b)) for a, b in p2}
    new1, new2 = [], []
    for q, t in p1:
        if (int(t), int(q)) in s2:
            new1.append([int(q), int(t)])
            new2.append([int(t), int(q)])
    return new1, new2


def filter_duplicates(pts1: np.ndarray, pts2: np.ndarray, idx_pairs: Sequence[Sequence[int]]):
    """matcher.py:157-182: first-seen-wins on '%.2f-%.2f' keypoint keys."""
    result = []
    d1, d2 = {}, {}
    for q, t in idx_pairs:
        k1 = "%.2f-%.2f" % (pts1[q][0], pts1[q][1])
        k2 = "%.2f-%.2f" % (pts2[t][0], pts2[t][1])
        if k1 in d1 or k2 in d2:
            continue
        d1[k1] = True
        d2[k2] = True
        result.append([int(q), int(t)])
    return result


def basic_pair(q: np.ndarray, t: np.ndarray, norm: int, match_ratio: float, max_distance: float, cap: int = 2000,
               min_pairs: int = 25, threads: int = 1, mode: str = "ref_metric", pts_q=None, pts_t=None, size=None,
               dedupe: bool = False, gms_kw=None):
    """basic_pair_matches (matcher.py:218-300) for one direction: kNN, metric reduction + gate (:220-273),
    then -- when `size` is given -- the GMS filter (:285), and -- with `dedupe` -- filter_duplicates and its
    gate (:294-298).  The reference gates once more only after filter_duplicates; without it the survivors of
    GMS are gated the same way (what the device pipeline does when no dedupe stage follows)."""
    i1, d1 = knn(q, t, 2, norm, threads)
    p = reduce_ref_metric(i1, d1, match_ratio, max_distance, cap, min_pairs) if mode == "ref_metric" \
        else reduce_lowe(i1, d1, match_ratio, cap, min_pairs)
    if size is not None and len(p) > 0:
        mask = gms_mask(pts_q, pts_t, size, size, p, **(gms_kw or {}))
        p = [m for m, keep in zip(p, mask) if keep]
    if dedupe:
        p = filter_duplicates(pts_q, pts_t, p)
    if (size is not None or dedupe) and len(p) < min_pairs:
        p = []
    return p


def bidirectional(q: np.ndarray, t: np.ndarray, norm: int, match_ratio: float, max_distance: float, cap: int = 2000,
                  min_pairs: int = 25, threads: int = 1, mode: str = "ref_metric", cross_check: bool = True,
                  pts_q=None, pts_t=None, size=None, dedupe: bool = False, gms_kw=None):
    """bidirectional_pair_matches (matcher.py:304-318): forward basic_pair_matches, reverse only if forward
    >= min_pairs, cross-check.  GMS runs when `size` (and the keypoint coordinates) are given,
    filter_duplicates with `dedupe`; by default both are off (descriptor-only callers)."""
    kw = dict(cap=cap, min_pairs=min_pairs, threads=threads, mode=mode, size=size, dedupe=dedupe, gms_kw=gms_kw)
    p1 = basic_pair(q, t, norm, match_ratio, max_distance, pts_q=pts_q, pts_t=pts_t, **kw)
    if not cross_check:
        return p1, None
    if len(p1) >= min_pairs and len(p1) > 0:
        p2 = basic_pair(t, q, norm, match_ratio, max_distance, pts_q=pts_t, pts_t=pts_q, **kw)
    else:
        p2 = []
    return filter_cross_check(p1, p2)


# --------------------------------------------------------------------------
# pair work-list  (matcher.py:858-903)
# --------------------------------------------------------------------------
def worklist(neds: np.ndarray, mode: str = "sequential", min_dist: float = 0.0, max_dist: float | None = None,
             seq_k: int = 4):
    """matcher.py:858-903.  `sequential` is the live branch (:899), `geotag`
    the documented distance window (:896, disabled by `if False` in the
    snapshot).  Returns [[ddist, i, j], ...] in generation order."""
    n = len(neds)
    intervals = [float(np.linalg.norm(np.array(neds[i + 1]) - np.array(neds[i]))) for i in range(n - 1)]
    median = float(np.median(intervals))
    average = float(np.average(intervals))
    if median < average:
        median = average
    median_int = int(round(median))
    if median_int == 0:
        median_int = 1
    if max_dist is None:
        max_dist = median_int * 4
    interval = median_int * 1.3
    work = []
    for i in range(n):
        for j in range(i + 1, n):
            dist = float(np.linalg.norm(np.array(neds[j]) - np.array(neds[i])))
            if mode == "geotag":
                if dist >= min_dist and dist <= max_dist:
                    work.append([int(round(dist / interval)) * interval, i, j])
            elif abs(i - j) <= seq_k:
                work.append([int(round(dist / interval)) * interval, i, j])
    return work


# --------------------------------------------------------------------------
# GMS grid-motion-statistics filter  (matcher.py:285 cv2.xfeatures2d.matchGMS;
# restated from the reference's own scripts/lib/archive/gms_matcher.py, which
# its header says reproduces the OpenCV C++ output)
# --------------------------------------------------------------------------
# Which right-cell neighbour faces left-cell neighbour j under each of the 8 rotations
# (gms_matcher.py:29-60; 1-based there, 0-based here).
GMS_ROTATIONS = [
    [0, 1, 2, 3, 4, 5, 6, 7, 8], [3, 0, 1, 6, 4, 2, 7, 8, 5], [6, 3, 0, 7, 4, 1, 8, 5, 2], [7, 6, 3, 8, 4, 0, 5, 2, 1],
    [8, 7, 6, 5, 4, 3, 2, 1, 0], [5, 8, 7, 2, 4, 6, 1, 0, 3], [2, 5, 8, 1, 4, 7, 0, 3, 6], [1, 2, 5, 0, 4, 8, 3, 6, 7]]
GMS_SCALES = [1.0, 0.5, 1.0 / np.sqrt(2.0), np.sqrt(2.0), 2.0]   # gms_matcher.py:63
GMS_GRID = 20                                                    # gms_matcher.py:83


def _gms_nb9(cell: int, gw: int, gh: int):
    """3x3 neighbourhood of a cell, -1 outside the grid (gms_matcher.py:112-127)."""
    cx, cy = cell % gw, cell // gw
    out = []
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            x, y = cx + dx, cy + dy
            out.append(x + y * gw if 0 <= x < gw and 0 <= y < gh else -1)
    return out


def _gms_run(lcell, rcell, n_right, gw_r, gh_r, rot, factor, archive_wrap=False):
    """One GmsMatcher.run(RotationType) (gms_matcher.py:187-209): the four half-cell shifted left grids,
    each with its own statistics, cell pairing and verification; a match is an inlier if any of the four accepts it.
    lcell: [4][n] left cell per grid type (-1 = outside); rcell: [n] right cell."""
    n = len(rcell)
    n_left = GMS_GRID * GMS_GRID
    mask = np.zeros(n, bool)
    for g in range(4):
        stats = np.zeros((n_left, n_right), np.int64)
        per_left = np.zeros(n_left, np.int64)
        for i in range(n):                                        # AssignMatchPairs :211-226
            lg, rg = lcell[g][i], rcell[i]
            if lg < 0 or rg < 0:
                continue
            stats[lg, rg] += 1
            per_left[lg] += 1
        pair = np.full(n_left, -1, np.int64)
        for i in range(n_left):                                   # VerifyCellPairs :253-285
            if stats[i].sum() == 0:
                continue
            j = int(np.argmax(stats[i]))                          # first maximum, as the strict '>' scan :261-265
            nl = _gms_nb9(i, GMS_GRID, GMS_GRID)
            nr = _gms_nb9(j, gw_r, gh_r)
            score = thresh = numpair = 0
            for s in range(9):
                ll, rr = nl[s], nr[GMS_ROTATIONS[rot][s]]
                if ll == -1 or rr == -1:
                    continue
                score += stats[ll, rr]
                thresh += per_left[ll]
                numpair += 1
            pair[i] = j if score >= factor * np.sqrt(thresh / numpair) else -2
        for i in range(n):                                        # mark inliers :203-207
            lc = lcell[g][i]
            if lc < 0:
                # A key point in the last half cell has no cell in a shifted grid (index -1).  OpenCV's C++
                # (gms.cpp run(): "if (mvMatchPairs[i].first >= 0)") skips the match for that grid; the archive
                # Python has no such test and reads mCellPairs[-1], i.e. wraps around to the LAST cell (:205).
                # archive_wrap=True restates the Python literally (tests pin it against the module); the default is
                # the C++ rule, which is what the call site (cv2.xfeatures2d.matchGMS, matcher.py:285) executes and
                # what the CUDA kernel does.
                if not archive_wrap:
                    continue
                lc = n_left - 1
            if pair[lc] == rcell[i]:
                mask[i] = True
    return mask


def gms_mask(pts1, pts2, size1, size2, matches, with_rotation: bool = True, with_scale: bool = False,
             threshold_factor: float = 5.0, archive_wrap: bool = False) -> np.ndarray:
    """Inlier mask over `matches` ([[queryIdx, trainIdx], ...]) as GmsMatcher.GetInlierMask returns it
    (gms_matcher.py:129-176): the rotation (and scale) hypothesis with the most inliers, first one on ties;
    the reference calls it with withRotation=True, withScale=False, thresholdFactor=5.0 (matcher.py:285).
    pts are pixel coordinates, size = (width, height) (gms_matcher.py:91-97)."""
    matches = np.asarray(matches, np.int64).reshape(-1, 2)
    n = len(matches)
    if n == 0:
        return np.zeros(0, bool)
    p1 = np.asarray(pts1, np.float64)[matches[:, 0]] / np.array(size1, np.float64)
    p2 = np.asarray(pts2, np.float64)[matches[:, 1]] / np.array(size2, np.float64)
    lcell = []
    for sx, sy in ((0.0, 0.0), (0.5, 0.0), (0.0, 0.5), (0.5, 0.5)):      # GetGridIndexLeft :228-246
        x = np.floor(p1[:, 0] * GMS_GRID + sx).astype(np.int64)
        y = np.floor(p1[:, 1] * GMS_GRID + sy).astype(np.int64)
        lcell.append(np.where((x >= GMS_GRID) | (y >= GMS_GRID), -1, x + y * GMS_GRID))
    best_mask, best = None, 0
    last = np.zeros(n, bool)
    for sc in (range(5) if with_scale else (0,)):
        gw = int(GMS_GRID * GMS_SCALES[sc])                              # SetScale :178-185
        gh = int(GMS_GRID * GMS_SCALES[sc])
        rcell = np.floor(p2[:, 0] * gw).astype(np.int64) + np.floor(p2[:, 1] * gh).astype(np.int64) * gw  # :248-251
        for rot in (range(8) if with_rotation else (0,)):
            last = _gms_run(lcell, rcell, gw * gh, gw, gh, rot, threshold_factor, archive_wrap)
            c = int(last.sum())
            if c > best:
                best, best_mask = c, last
    return best_mask if best_mask is not None else last


# --------------------------------------------------------------------------
# bundle-adjustment residual  (scripts/lib/optimizer.py:174-279 Optimizer.fun,
# cam_method 'ned_quat' :84-85; cv2.projectPoints restated from the OpenCV
# pinhole + 5-coefficient distortion model)
# --------------------------------------------------------------------------
BA_CAM2BODY = np.array([[0.0, 0.0, 1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])   # optimizer.py:91-93


def ba_rotation(quat) -> np.ndarray:
    """ned -> camera rotation of nedquat2rvectvec (optimizer.py:121-127): quaternion_matrix normalises the
    (w, x, y, z) quaternion (transformations.py quaternion_matrix), body2ned = that matrix, R = body2cam . ned2body.
    (The reference then passes R through cv2.Rodrigues and back inside projectPoints: the same rotation.)"""
    q = np.asarray(quat, np.float64)
    n = float(q @ q)
    if n < np.finfo(float).eps * 4.0:
        b2n = np.identity(3)
    else:
        w, x, y, z = q * np.sqrt(2.0 / n)           # outer(q, q) below then carries the factor 2
        b2n = np.array([[1.0 - y * y - z * z, x * y - z * w, x * z + y * w],
                        [x * y + z * w, 1.0 - x * x - z * z, y * z - x * w],
                        [x * z - y * w, y * z + x * w, 1.0 - x * x - y * y]])
    return np.linalg.inv(BA_CAM2BODY) @ b2n.T


def ba_project(X: np.ndarray, R: np.ndarray, ned: np.ndarray, K4, dist) -> np.ndarray:
    """cv2.projectPoints(X, rvec(R), tvec = -R ned, K, distCoeffs) (optimizer.py:220): [n, 2] pixels."""
    fx, fy, cx, cy = K4
    k1, k2, p1, p2, k3 = dist
    Xc = (np.asarray(X, np.float64) - np.asarray(ned, np.float64)) @ R.T
    x, y = Xc[:, 0] / Xc[:, 2], Xc[:, 1] / Xc[:, 2]
    r2 = x * x + y * y
    rad = 1.0 + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2
    xd = x * rad + 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x)
    yd = y * rad + p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y
    return np.stack([fx * xd + cx, fy * yd + cy], 1)


def ba_residuals(params, n_cam: int, n_pts: int, cam_idx, pt_idx, obs_uv, K4, dist) -> np.ndarray:
    """Optimizer.fun (optimizer.py:174-279) for optimize_calib 'none': observed - projected, (u, v) interleaved,
    observations grouped by camera in list order (`cam_idx` non-decreasing, as :400-404 lays them out)."""
    params = np.asarray(params, np.float64)
    cams = params[:n_cam * 7].reshape(n_cam, 7)
    pts = params[n_cam * 7:n_cam * 7 + n_pts * 3].reshape(n_pts, 3)
    cam_idx = np.asarray(cam_idx)
    out = np.zeros((len(cam_idx), 2))
    for c in np.unique(cam_idx):
        sel = np.flatnonzero(cam_idx == c)
        R = ba_rotation(cams[c, 3:7])
        out[sel] = np.asarray(obs_uv, np.float64)[sel] - ba_project(pts[np.asarray(pt_idx)[sel]], R, cams[c, :3], K4, dist)
    return out.ravel()


def ba_jacobian_fd(params, n_cam, n_pts, cam_idx, pt_idx, obs_uv, K4, dist, rel_step: float = 1e-6) -> np.ndarray:
    """Central-difference Jacobian blocks of ba_residuals, [n_obs][2][10]: columns 0..6 = the observation's camera
    parameters (ned, quat), 7..9 = its 3-D point.  Checker for the analytic Jacobian kernel (the reference itself
    lets SciPy difference fun(): least_squares(..., jac_sparsity=A), optimizer.py:491-501)."""
    params = np.asarray(params, np.float64)
    n_obs = len(cam_idx)
    J = np.zeros((n_obs, 2, 10))
    cam_idx, pt_idx = np.asarray(cam_idx), np.asarray(pt_idx)
    for col in range(10):
        # one column of every camera (or point) block at a time: blocks of different cameras/points do not interact
        hp = np.zeros_like(params)
        if col < 7:
            cols = np.arange(n_cam) * 7 + col
        else:
            cols = n_cam * 7 + np.arange(n_pts) * 3 + (col - 7)
        h = rel_step * np.maximum(1.0, np.abs(params[cols]))
        hp[cols] = h
        f1 = ba_residuals(params + hp, n_cam, n_pts, cam_idx, pt_idx, obs_uv, K4, dist).reshape(n_obs, 2)
        f0 = ba_residuals(params - hp, n_cam, n_pts, cam_idx, pt_idx, obs_uv, K4, dist).reshape(n_obs, 2)
        step = h[cam_idx] if col < 7 else h[pt_idx]
        J[:, :, col] = (f1 - f0) / (2.0 * step[:, None])
    return J


# --------------------------------------------------------------------------
# per-query neighbour selection of the bin-fitting strategies, restated literally
# (scripts/lib/matcher.py:484-514 smart_pair_matches, :709-753 bruteforce_pair_matches)
# --------------------------------------------------------------------------
def best_of_neighbours_literal(idx, dist, pts1, size1, pts2, size2, match_ratio, dist_limit, pred1=None):
    """For every query row walk its k neighbours exactly as the reference's Python loops do.  pred1 = predicted
    positions of the image-1 key points in image 2 (smart: metric = raw_dist * size_diff / ratio, :505); None =
    bruteforce (metric = size_diff / ratio, :744; raw_dist = |p2 - p1| in float32, :725-728).
    Returns [(row, best_j, raw_dist, vangle), ...] for rows with a surviving neighbour."""
    import math
    out = []
    for i in range(idx.shape[0]):
        best_index, best_metric, best_dist, best_vangle = -1, 9, 0.0, 0.0
        for j in range(idx.shape[1]):
            if idx[i, j] < 0:
                break
            if dist[i, j] >= dist_limit:                      # :491 / :718
                break
            if float(dist[i, j]) == 0.0 and float(dist[i, 0]) == 0.0:
                break                                         # the reference would raise ZeroDivisionError here
            with np.errstate(divide="ignore"):
                ratio = float(dist[i, 0]) / float(dist[i, j]) if dist[i, j] != 0 else float("inf")
            if ratio < match_ratio:                           # :494 / :721
                break
            t = int(idx[i, j])
            p1 = np.float32(pts1[i]) if pred1 is None else np.float32(pred1[i])
            p2 = np.float32(pts2[t])
            v = p2 - p1
            raw_dist = float(np.linalg.norm(v))
            vangle = math.atan2(float(v[1]), float(v[0]))
            if vangle < 0:
                vangle += 2 * math.pi
            s1, s2 = float(size1[i]), float(size2[t])
            size_diff = s1 / s2 if s1 > s2 else s2 / s1       # :499-502 / :736-739
            if size_diff > 1.25:
                continue
            metric = (raw_dist * size_diff / ratio) if pred1 is not None else (size_diff / ratio)
            if best_index < 0 or metric < best_metric:
                best_metric, best_index, best_dist, best_vangle = metric, j, raw_dist, vangle
        if best_index >= 0:
            out.append((i, best_index, best_dist, best_vangle))
    return out
