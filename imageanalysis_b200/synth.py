"""Synthetic inputs of the shapes BASELINE.json names.

There is no network and no survey imagery in the build environment, so the
benchmark and the tests run on descriptors with the statistics of OpenCV's
output: SIFT rows are integer valued in [0,255] with L2 norm ~512 (what
cv2.SIFT_create() returns; reference image.py:237,324), ORB rows are 32
random bytes; correspondences between neighbouring frames are planted by
copying rows with small perturbations.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def sift_like(n: int, seed: int, dim: int = 128) -> np.ndarray:
    """[n, dim] uint8: gamma(0.6) magnitudes, L2-normalise, clamp 0.2,
    renormalise, x512, round, clip — the post-processing chain of Lowe's
    descriptor, which is what gives real SIFT its value histogram."""
    rng = np.random.default_rng(seed)
    v = rng.gamma(0.6, 1.0, (n, dim)).astype(np.float32)
    v /= np.maximum(np.linalg.norm(v, axis=1, keepdims=True), 1e-12)
    v = np.minimum(v, 0.2)
    v /= np.maximum(np.linalg.norm(v, axis=1, keepdims=True), 1e-12)
    return np.clip(np.rint(v * 512.0), 0, 255).astype(np.uint8)


def orb_like(n: int, seed: int, nbytes: int = 32) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (n, nbytes), dtype=np.uint8)


def flip_bits(rows: np.ndarray, max_flips: int, rng: np.random.Generator) -> np.ndarray:
    out = rows.copy()
    nbits = rows.shape[1] * 8
    for r in range(out.shape[0]):
        k = int(rng.integers(0, max_flips + 1))
        for b in rng.choice(nbits, size=k, replace=False):
            out[r, b >> 3] ^= np.uint8(1 << (b & 7))
    return out


def plant(prev: np.ndarray, cur: np.ndarray, frac: float, rng: np.random.Generator, kind: str):
    """Overwrite a random `frac` of cur's rows with perturbed copies of random
    rows of prev.  Returns (dst_rows, src_rows)."""
    n = cur.shape[0]
    m = int(round(frac * min(n, prev.shape[0])))
    src = rng.permutation(prev.shape[0])[:m]
    dst = rng.permutation(n)[:m]
    if kind == "sift":
        noise = rng.integers(-3, 4, (m, cur.shape[1]))
        cur[dst] = np.clip(prev[src].astype(np.int32) + noise, 0, 255).astype(np.uint8)
    else:
        cur[dst] = flip_bits(prev[src], 20, rng)
    return dst, src


def sift_project(n_images: int, n_desc: int, seed: int = 0, planted: float = 0.4, kind: str = "sift"):
    """A strip of n_images frames: descriptors [n_desc, D] uint8, keypoint pixel
    coordinates (5472x3648 frame, cameras/DJI_FC6310S.json) and NED camera
    positions 15 m apart.  Frame i shares `planted` of its rows with frame i-1."""
    rng = np.random.default_rng(seed)
    des: List[np.ndarray] = []
    pts: List[np.ndarray] = []
    neds: List[List[float]] = []
    shift = np.float32([-733.0, 12.0])  # 15 m at 75 m AGL with f=3666 px
    for i in range(n_images):
        d = sift_like(n_desc, seed * 100003 + i) if kind == "sift" else orb_like(n_desc, seed * 100003 + i)
        p = np.stack([rng.uniform(0, 5472, n_desc), rng.uniform(0, 3648, n_desc)], 1).astype(np.float32)
        if i > 0:
            dst, src = plant(des[i - 1], d, planted, rng, kind)
            p[dst] = pts[i - 1][src] + shift + rng.normal(0, 0.3, (len(dst), 2)).astype(np.float32)
        des.append(d)
        pts.append(p)
        neds.append([15.0 * i, 0.0, -75.0])
    return des, pts, neds


def survey_grid_neds(lines: int = 38, per_line: int = 74, along: float = 15.0, cross: float = 25.0,
                     agl: float = 75.0) -> np.ndarray:
    """The 2812-frame Bates-survey shape (reference README.md:26-27) as a
    serpentine lawn-mower grid: 38 lines x 74 frames."""
    out = []
    for l in range(lines):
        xs = range(per_line) if l % 2 == 0 else range(per_line - 1, -1, -1)
        for x in xs:
            out.append([x * along, l * cross, -agl])
    return np.asarray(out, np.float64)


def two_view_scene(n: int, outlier_frac: float, K: np.ndarray, seed: int = 0, noise_px: float = 0.5,
                   width: int = 5472, height: int = 3648) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Pixel correspondences of a nadir camera pair over rough terrain plus a
    fraction of uniformly random outliers.  Returns (p1 [n,2] f32, p2 [n,2] f32,
    truth [n] u8 with 1 = planted inlier)."""
    rng = np.random.default_rng(seed)
    X = np.stack([rng.uniform(-45, 45, n), rng.uniform(-30, 30, n), rng.uniform(60, 90, n)], 1)
    ang = np.deg2rad(rng.uniform(2, 6))
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx
    t = np.array([-15.0, 1.0, 0.5])

    def proj(Xc):
        x = (K @ Xc.T).T
        return x[:, :2] / x[:, 2:3]

    p1 = proj(X) + rng.normal(0, noise_px, (n, 2))
    p2 = proj((R @ X.T).T + t) + rng.normal(0, noise_px, (n, 2))
    truth = np.ones(n, np.uint8)
    n_out = int(round(outlier_frac * n))
    bad = rng.permutation(n)[:n_out]
    p2[bad] = np.stack([rng.uniform(0, width, n_out), rng.uniform(0, height, n_out)], 1)
    truth[bad] = 0
    return p1.astype(np.float32), p2.astype(np.float32), truth


def ba_problem(n_cam: int, n_pts: int, seed: int = 0, obs_per_cam: int = 200, empty=(), width: int = 5472,
               height: int = 3648):
    """A bundle-adjustment problem of the shape the reference's Optimizer.setup() builds (optimizer.py:283-420):
    nadir cameras on a strip 15 m apart at 75 m AGL as [ned(3), quat(4)] rows (cam_method 'ned_quat', :84-85),
    3-D points on rough ground, per camera a list of point indices it observes and their (distorted) pixel
    coordinates plus noise.  Returns a dict: params (cameras then points, as x0 :422-423), n_cam, n_pts,
    idx_lists, uv_lists, K, dist."""
    rng = np.random.default_rng(seed)
    K = np.array([[3666.666504, 0.0, width / 2.0], [0.0, 3666.666504, height / 2.0], [0.0, 0.0, 1.0]])
    dist = np.array([-0.012, 0.006, 0.0004, -0.0002, 0.001])
    cam2body = np.array([[0.0, 0, 1], [1, 0, 0], [0, 1, 0]])
    cams = np.zeros((n_cam, 7))
    for i in range(n_cam):
        yaw, pitch, roll = np.deg2rad(rng.normal(0, 3.0)), np.deg2rad(-90 + rng.normal(0, 2.0)), np.deg2rad(rng.normal(0, 2.0))
        cy, sy, cp, sp, cr, sr = np.cos(yaw / 2), np.sin(yaw / 2), np.cos(pitch / 2), np.sin(pitch / 2), np.cos(roll / 2), np.sin(roll / 2)
        # body -> ned quaternion (w, x, y, z) of a ZYX euler sequence; deliberately NOT unit length: the reference
        # normalises inside quaternion_matrix and the optimiser is free to scale it
        q = np.array([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy])
        cams[i, :3] = [15.0 * i + rng.normal(0, 0.5), rng.normal(0, 0.5), -75.0 + rng.normal(0, 0.5)]
        cams[i, 3:] = q * rng.uniform(0.8, 1.25)
    pts = np.stack([rng.uniform(-60, 15.0 * n_cam + 60, n_pts), rng.uniform(-45, 45, n_pts), rng.normal(0, 3.0, n_pts)], 1)
    idx_lists, uv_lists = [], []
    order = np.argsort(pts[:, 0], kind="stable")      # only points near the camera's footprint are projected
    xs = pts[order, 0]
    for i in range(n_cam):
        if i in empty:   # shapes as optimizer.py:390-392 leaves them
            idx_lists.append(np.array([], np.int64))
            uv_lists.append(np.zeros((0, 1, 2)))
            continue
        w, x, y, z = cams[i, 3:] / np.linalg.norm(cams[i, 3:])
        b2n = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        R = cam2body.T @ b2n.T
        near = np.sort(order[np.searchsorted(xs, cams[i, 0] - 90.0):np.searchsorted(xs, cams[i, 0] + 90.0)])
        Xc = (pts[near] - cams[i, :3]) @ R.T
        xn, yn = Xc[:, 0] / Xc[:, 2], Xc[:, 1] / Xc[:, 2]
        r2 = xn * xn + yn * yn
        rad = 1 + dist[0] * r2 + dist[1] * r2 * r2 + dist[4] * r2 ** 3
        xd = xn * rad + 2 * dist[2] * xn * yn + dist[3] * (r2 + 2 * xn * xn)
        yd = yn * rad + dist[2] * (r2 + 2 * yn * yn) + 2 * dist[3] * xn * yn
        u, v = K[0, 0] * xd + K[0, 2], K[1, 1] * yd + K[1, 2]
        vis = np.flatnonzero((Xc[:, 2] > 1) & (u >= 0) & (u < width) & (v >= 0) & (v < height))
        pick = rng.permutation(vis)[:obs_per_cam]
        idx_lists.append(np.asarray(near[pick], np.int64))
        uv_lists.append((np.stack([u[pick], v[pick]], 1) + rng.normal(0, 1.5, (len(pick), 2))).reshape(len(pick), 1, 2))
    params = np.concatenate([cams.ravel(), (pts + rng.normal(0, 0.3, pts.shape)).ravel()])
    return dict(params=params, n_cam=n_cam, n_pts=n_pts, idx_lists=idx_lists, uv_lists=uv_lists, K=K, dist=dist)
