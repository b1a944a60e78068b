#!/usr/bin/env python3
"""Golden fixtures for the alternative matching strategies and the robust-fit siblings:

  * the reference's own `lib.matcher` (imported UNMODIFIED from /root/reference, shims in ./shims, exact
    cv2.BFMatcher injected in place of FLANN) running
        smart_pair_matches      (matcher.py:358-593)   -- pose-predicted homography, k = 3, bins by predicted error
        ratio_pair_matches      (matcher.py:595-694)   -- bins by Lowe ratio
        bruteforce_pair_matches (matcher.py:696-850)   -- bins by displacement length x direction
    on a synthetic two-frame scene with real camera poses (nadir cameras over flat ground, so a homography between
    the frames exists), key points with sizes and angles, and planted descriptor correspondences;
  * the preliminary homography smart_pair_matches fits through the 81 re-projected grid points (:452), captured from
    the reference's own cv2.findHomography call;
  * cv2.findHomography(src, dst, cv2.RANSAC, tol) and cv2.findFundamentalMat(p1, p2, cv2.RANSAC, tol) exactly as
    matcher.py:122-124 calls them, on synthetic scenes.

The homography / fundamental RANSACs cannot be bit-identical between OpenCV and the CUDA kernel (different
samplers): tests compare inlier SETS with a stated tolerance.

usage: python tests/golden/make_golden_strategies.py      (from the repo root; needs /root/reference and cv2)
"""
import contextlib
import io
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, os.path.join(REF, "scripts", "lib", "archive"))
sys.path.insert(0, os.path.join(REF, "scripts"))

import cv2  # noqa: E402
from transformations import quaternion_from_euler, quaternion_matrix, rotation_matrix  # noqa: E402

from imageanalysis_b200 import synth  # noqa: E402

W, H = 5472, 3648
K = np.array([[3666.666504, 0.0, 2736.0], [0.0, 3666.666504, 1824.0], [0.0, 0.0, 1.0]])
CAM2BODY = np.array([[0.0, 0.0, 1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])      # optical axis = body x (optimizer.py:91-93)
d2r = math.pi / 180.0


class PoseImage:
    """Duck-typed lib.image.Image with the pose API smart_pair_matches touches (image.py:507-553)."""

    def __init__(self, name, des, pts, sizes, angles, ned, yaw_deg):
        self.name = name
        self.des_list = des
        self.kp_list = [cv2.KeyPoint(x=float(p[0]), y=float(p[1]), size=float(s), angle=float(a))
                        for p, s, a in zip(pts, sizes, angles)]
        self.uv_list = [list(map(float, p)) for p in pts]
        self.match_list = {}
        self.ned = [float(v) for v in ned]
        # camera pose = aircraft yaw with the camera mount pitched straight down (ypr = yaw, -90, 0): with the base
        # cam2body below the optical axis is the body x axis, which then points along +D
        self.quat = quaternion_from_euler(yaw_deg * d2r, -90.0 * d2r, 0.0, "rzyx")
        self.cam2body = CAM2BODY
        self.body2cam = CAM2BODY.T

    def get_camera_pose(self, opt=False):
        return self.ned, [0.0, 0.0, 0.0], list(self.quat)

    def get_cam2body(self):
        return self.cam2body

    def get_body2cam(self):
        return self.body2cam

    def get_body2ned(self, opt=False):                      # image.py:537-539
        return quaternion_matrix(np.array(self.quat))[:3, :3]

    def get_ned2body(self, opt=False):                      # image.py:533-534
        return np.matrix(self.get_body2ned(opt)).T

    def get_proj(self, opt=False, yaw_error_est=0.0):       # image.py:543-553
        body2cam = self.get_body2cam()
        ned2body = self.get_ned2body(opt)
        if abs(yaw_error_est) > 0.001 and not opt:
            R1 = rotation_matrix(yaw_error_est * d2r, [1, 0, 0])[:3, :3]
            ned2body = np.dot(self.get_body2ned(), R1).T
        R = body2cam.dot(ned2body)
        rvec, _ = cv2.Rodrigues(R)
        tvec = -np.matrix(R) * np.matrix(self.ned).T
        return rvec, tvec


def project(img, X):
    """Pixels of NED points X in img (the model Image.get_proj + cv2.projectPoints implement, zero distortion)."""
    R = np.asarray(img.get_body2cam().dot(np.asarray(img.get_ned2body())))
    Xc = (np.asarray(X) - np.asarray(img.ned)) @ R.T
    uv = Xc @ K.T
    return uv[:, :2] / uv[:, 2:3]


def scene(seed, n=1800, planted=0.45):
    """Two nadir frames 18 m apart (75 m above flat ground, 4 degrees of relative yaw)."""
    rng = np.random.default_rng(seed)
    a = PoseImage("A", None, np.zeros((0, 2)), [], [], (0.0, 0.0, -75.0), 0.0)
    b = PoseImage("B", None, np.zeros((0, 2)), [], [], (18.0, 3.0, -75.0), 4.0)
    # ground points seen by both (flat ground at D = 0)
    m = int(planted * n)
    X = np.stack([rng.uniform(-30, 50, 4 * m), rng.uniform(-50, 50, 4 * m), np.zeros(4 * m)], 1)
    pa, pb = project(a, X), project(b, X)
    ok = (pa[:, 0] > 0) & (pa[:, 0] < W) & (pa[:, 1] > 0) & (pa[:, 1] < H) & \
         (pb[:, 0] > 0) & (pb[:, 0] < W) & (pb[:, 1] > 0) & (pb[:, 1] < H)
    pa, pb = pa[ok][:m], pb[ok][:m]
    m = len(pa)
    pts_a = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], 1)
    pts_b = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], 1)
    ia, ib = rng.permutation(n)[:m], rng.permutation(n)[:m]
    pts_a[ia] = pa + rng.normal(0, 0.4, (m, 2))
    pts_b[ib] = pb + rng.normal(0, 0.4, (m, 2))
    size_a = rng.uniform(2.0, 9.0, n)
    size_b = rng.uniform(2.0, 9.0, n)
    size_b[ib] = size_a[ia] * rng.uniform(0.9, 1.12, m)          # most planted pairs pass the 1.25 size gate
    size_b[ib[: m // 12]] = size_a[ia[: m // 12]] * 1.4          # ... some do not
    ang_a, ang_b = rng.uniform(0, 360, n), rng.uniform(0, 360, n)
    ang_b[ib] = (ang_a[ia] + 4.0 + rng.normal(0, 2.0, m)) % 360
    des_a = synth.sift_like(n, seed=seed * 10 + 1)
    des_b = synth.sift_like(n, seed=seed * 10 + 2)
    des_b[ib] = np.clip(des_a[ia].astype(np.int32) + rng.integers(-4, 5, (m, 128)), 0, 255).astype(np.uint8)
    # a few ambiguous rows: the planted row and a near copy of it elsewhere (second/third neighbours matter)
    dup = ib[: m // 10]
    other = rng.permutation(np.setdiff1d(np.arange(n), ib))[: len(dup)]
    des_b[other] = np.clip(des_b[dup].astype(np.int32) + rng.integers(-9, 10, (len(dup), 128)), 0, 255).astype(np.uint8)
    a = PoseImage("A", des_a.astype(np.float32), pts_a.astype(np.float32), size_a, ang_a, a.ned, 0.0)
    b = PoseImage("B", des_b.astype(np.float32), pts_b.astype(np.float32), size_b, ang_b, b.ned, 4.0)
    return a, b, dict(des_a=des_a, des_b=des_b, pts_a=pts_a.astype(np.float32), pts_b=pts_b.astype(np.float32),
                      size_a=size_a, size_b=size_b, ang_a=ang_a, ang_b=ang_b, ned_a=np.float64(a.ned),
                      ned_b=np.float64(b.ned), quat_a=np.float64(a.quat), quat_b=np.float64(b.quat),
                      planted_a=ia, planted_b=ib)


def import_reference_matcher():
    from props import getNode
    det = getNode("/config/detector", True)
    det.setString("detector", "SIFT")
    det.setFloat("scale", 1.0)
    mn = getNode("/config/matcher", True)
    mn.setFloat("match_ratio", 0.75)
    mn.setFloat("min_pairs", 25)
    mn.setFloat("ground_m", 0.0)                   # "Forced ground" (matcher.py:372-374): no SRTM needed
    cam = getNode("/config/camera", True)
    cam.setInt("width_px", W)
    cam.setInt("height_px", H)
    cam.setLen("K", 9, 0.0)
    for i, v in enumerate(K.ravel()):
        cam.setFloatEnum("K", i, float(v))
    cam.setLen("dist_coeffs", 5, 0.0)
    from lib import matcher
    matcher.configure()
    matcher.the_matcher = cv2.BFMatcher(cv2.NORM_L2)
    return matcher


def gen_strategies():
    matcher = import_reference_matcher()
    out = {"K": K, "size": np.int32([W, H])}
    for s, seed in enumerate((3, 4)):
        a, b, data = scene(seed)
        for k, v in data.items():
            out["s%d_%s" % (s, k)] = v
        captured = {}
        real_fh = cv2.findHomography

        def spy(src, dst, method=0, *args, **kw):
            r = real_fh(src, dst, method, *args, **kw)
            if method == 0 and "H0" not in captured:
                captured["H0"] = np.array(r[0])
            return r

        for name, fn, kw in (("smart", matcher.smart_pair_matches, dict(review=False, est_rotation=False)),
                             ("ratio", matcher.ratio_pair_matches, dict(review=False, est_rotation=False)),
                             ("bruteforce", matcher.bruteforce_pair_matches, dict(review=False))):
            cv2.findHomography = spy
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    fwd, rev = fn(a, b, **kw)
            finally:
                cv2.findHomography = real_fh
            assert [[p[1], p[0]] for p in fwd] == rev
            out["s%d_%s" % (s, name)] = np.int32(fwd).reshape(-1, 2)
            planted = set(zip(data["planted_a"].tolist(), data["planted_b"].tolist()))
            good = sum((int(q), int(t)) in planted for q, t in fwd)
            print("scene %d %-10s -> %4d pairs (%d of them planted)" % (s, name, len(fwd), good))
        out["s%d_H0" % s] = captured["H0"]
    np.savez_compressed(os.path.join(HERE, "reference_strategies.npz"), **out)


def gen_robust_fits():
    """cv2.findHomography / cv2.findFundamentalMat exactly as matcher.py:122-124 calls them."""
    tol = max(1.0, W ** 0.25)
    out = {"tol": tol, "K": K}
    rng = np.random.default_rng(11)
    for s, (n, frac) in enumerate([(1500, 0.35), (300, 0.5), (60, 0.15)]):
        # planar scene: a homography exists between the two views
        a, b, _ = scene(20 + s, n=max(2 * n, 200))
        X = np.stack([rng.uniform(-25, 45, 6 * n), rng.uniform(-45, 45, 6 * n), np.zeros(6 * n)], 1)
        pa, pb = project(a, X), project(b, X)
        ok = (pa > 0).all(1) & (pb > 0).all(1) & (pa[:, 0] < W) & (pb[:, 0] < W) & (pa[:, 1] < H) & (pb[:, 1] < H)
        p1 = (pa[ok][:n] + rng.normal(0, 0.5, (n, 2))).astype(np.float32)
        p2 = (pb[ok][:n] + rng.normal(0, 0.5, (n, 2))).astype(np.float32)
        truth = np.ones(n, np.uint8)
        bad = rng.permutation(n)[: int(frac * n)]
        p2[bad] = np.stack([rng.uniform(0, W, len(bad)), rng.uniform(0, H, len(bad))], 1)
        truth[bad] = 0
        Hm, mask = cv2.findHomography(p1, p2, cv2.RANSAC, tol)
        out["h_p1_%d" % s], out["h_p2_%d" % s], out["h_truth_%d" % s] = p1, p2, truth
        out["h_H_%d" % s], out["h_mask_%d" % s] = Hm, mask.ravel().astype(np.uint8)
        print("homography  scene %d: %4d points, %4d cv2 inliers, %4d planted" % (s, n, int(mask.sum()), int(truth.sum())))
        # general (non-planar) scene for the fundamental matrix
        q1, q2, qt = synth.two_view_scene(n, frac, K, seed=300 + s)
        F, fmask = cv2.findFundamentalMat(q1, q2, cv2.RANSAC, tol)
        out["f_p1_%d" % s], out["f_p2_%d" % s], out["f_truth_%d" % s] = q1, q2, qt
        out["f_F_%d" % s], out["f_mask_%d" % s] = F[:3], fmask.ravel().astype(np.uint8)
        print("fundamental scene %d: %4d points, %4d cv2 inliers, %4d planted" % (s, n, int(fmask.sum()), int(qt.sum())))
    np.savez_compressed(os.path.join(HERE, "robust_fits.npz"), **out)


if __name__ == "__main__":
    gen_strategies()
    gen_robust_fits()
