#!/usr/bin/env python3
"""Golden fixtures for the pair-wise side estimators of the 'smart' strategy (reference scripts/lib/smart.py), recorded by
running the reference's own `lib.smart` (imported UNMODIFIED from /root/reference, shims in ./shims) on synthetic posed
image pairs over rolling terrain:

    smart.triangulate_features(i1, i2)          (:26-63)    -> cv2.triangulatePoints
    smart.estimate_surface_elevation(i1, i2)    (:116-131)
    smart.find_affine(i1, i2)                   (:66-90)    -> cv2.estimateAffinePartial2D
    smart.estimate_yaw_error(i1, i2)            (:139-190)
    smart.update_surface_estimate / update_yaw_error_estimate / get_* over a 3-image chain (:194-317)

smart_reference.npz holds the scenes (poses, key points, match lists) and the reference's answers.
usage: python tests/golden/make_golden_smart.py      (from the repo root; needs /root/reference and cv2)
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_strategies as G  # noqa: E402  (sets sys.path for the reference + shims)

import cv2  # noqa: E402


YAW_BIAS = 3.0


class PoseImage(G.PoseImage):
    def __init__(self, name, pts, ned, yaw_deg):
        n = len(pts)
        super().__init__(name, None, pts, np.full(n, 3.0), np.zeros(n), ned, yaw_deg)
        self.yaw_deg = float(yaw_deg)
        self.yaw_error = 0.0

    def get_aircraft_pose(self):                      # image.py:494-505; the EKF yaw is off by YAW_BIAS degrees
        return [0.0, 0.0, 0.0], [self.yaw_deg - YAW_BIAS, 0.0, 0.0], [1.0, 0.0, 0.0, 0.0]

    def set_aircraft_yaw_error_estimate(self, yaw_error_deg):
        self.yaw_error = yaw_error_deg


def terrain(xy, seed, relief):
    rng = np.random.default_rng(seed)
    a, b, c = rng.uniform(-1, 1, 3)
    return 20.0 + relief * (np.sin(xy[:, 0] / 17.0 + a) * np.cos(xy[:, 1] / 23.0 + b) + 0.0125 * c * xy[:, 0])   # elevation, m


def make_pair(seed, ned_a, yaw_a, ned_b, yaw_b, n=900, outliers=0.0, noise=0.35, relief=4.0):
    rng = np.random.default_rng(seed)
    a0 = PoseImage("A", np.zeros((0, 2)), ned_a, yaw_a)
    b0 = PoseImage("B", np.zeros((0, 2)), ned_b, yaw_b)
    xy = np.stack([rng.uniform(-40, 60, 6 * n), rng.uniform(-60, 60, 6 * n)], 1)
    X = np.concatenate([xy, -terrain(xy, seed, relief)[:, None]], 1)          # NED: down = -elevation
    pa, pb = G.project(a0, X), G.project(b0, X)
    ok = (pa[:, 0] > 0) & (pa[:, 0] < G.W) & (pa[:, 1] > 0) & (pa[:, 1] < G.H) & \
         (pb[:, 0] > 0) & (pb[:, 0] < G.W) & (pb[:, 1] > 0) & (pb[:, 1] < G.H)
    pa, pb = pa[ok][:n], pb[ok][:n]
    n = len(pa)
    pa = pa + rng.normal(0, noise, pa.shape)
    pb = pb + rng.normal(0, noise, pb.shape)
    bad = rng.permutation(n)[: int(outliers * n)]
    pb[bad] = np.stack([rng.uniform(0, G.W, len(bad)), rng.uniform(0, G.H, len(bad))], 1)
    # key point lists hold extra, unmatched points; the match list indexes into them in a scrambled order
    extra = 150
    pts_a = np.concatenate([pa, np.stack([rng.uniform(0, G.W, extra), rng.uniform(0, G.H, extra)], 1)])
    pts_b = np.concatenate([pb, np.stack([rng.uniform(0, G.W, extra), rng.uniform(0, G.H, extra)], 1)])
    perm_a, perm_b = rng.permutation(len(pts_a)), rng.permutation(len(pts_b))
    inv_a, inv_b = np.argsort(perm_a), np.argsort(perm_b)
    a = PoseImage("A%d" % seed, pts_a[perm_a].astype(np.float32), ned_a, yaw_a)
    b = PoseImage("B%d" % seed, pts_b[perm_b].astype(np.float32), ned_b, yaw_b)
    matches = [[int(inv_a[i]), int(inv_b[i])] for i in rng.permutation(n)]
    a.match_list[b.name] = matches
    b.match_list[a.name] = [[m[1], m[0]] for m in matches]
    return a, b


def main():
    G.import_reference_matcher()            # fills /config/camera (K, size) in the shim property tree
    from lib import smart
    out = {"K": G.K, "size": np.int32([G.W, G.H])}
    # (seed, pose A, yaw A, pose B, yaw B, outlier fraction, terrain relief in m)
    scenes = [(21, (0.0, 0.0, -95.0), 0.0, (22.0, 4.0, -96.0), 6.0, 0.0, 4.0),
              (22, (5.0, -3.0, -120.0), 40.0, (-14.0, 12.0, -118.0), 33.0, 0.02, 4.0),
              (23, (0.0, 0.0, -80.0), -10.0, (3.0, 25.0, -80.5), -14.0, 0.2, 0.0)]
    for s, (seed, ned_a, yaw_a, ned_b, yaw_b, outl, relief) in enumerate(scenes):
        a, b = make_pair(seed, ned_a, yaw_a, ned_b, yaw_b, outliers=outl, relief=relief)
        with contextlib.redirect_stdout(io.StringIO()):
            pts = smart.triangulate_features(a, b)
            surf = smart.estimate_surface_elevation(a, b)
            aff = smart.find_affine(a, b)
            yaw = smart.estimate_yaw_error(a, b)
            yaw_rev = smart.estimate_yaw_error(b, a)
        uv1 = np.float32([[a.kp_list[m[0]].pt for m in a.match_list[b.name]]])
        uv2 = np.float32([[b.kp_list[m[1]].pt for m in a.match_list[b.name]]])
        _, status = cv2.estimateAffinePartial2D(uv2, uv1)
        pre = "s%d_" % s
        out.update({pre + "pts_a": np.float32([k.pt for k in a.kp_list]), pre + "pts_b": np.float32([k.pt for k in b.kp_list]),
                    pre + "matches": np.int32(a.match_list[b.name]), pre + "ned_a": np.float64(a.ned), pre + "ned_b": np.float64(b.ned),
                    pre + "yaw_a": a.yaw_deg, pre + "yaw_b": b.yaw_deg,
                    pre + "proj_a": np.concatenate([cv2.Rodrigues(a.get_proj()[0])[0], np.asarray(a.get_proj()[1])], 1),
                    pre + "proj_b": np.concatenate([cv2.Rodrigues(b.get_proj()[0])[0], np.asarray(b.get_proj()[1])], 1),
                    pre + "points": np.asarray(pts, np.float64), pre + "surface": np.float64(surf),
                    pre + "affine": np.asarray(aff, np.float64), pre + "affine_inliers": status.ravel().astype(np.uint8),
                    pre + "yaw": np.float64(yaw), pre + "yaw_rev": np.float64(yaw_rev)})
        print("scene %d: %d matches, surface %.2f m (std %.2f), affine rot %.3f deg, yaw error %.2f / %.2f deg, %d affine inliers" % (
            s, len(a.match_list[b.name]), surf[0], surf[1], smart.decompose_affine(aff)[0], yaw[0], yaw_rev[0], int(status.sum())))
    # the running estimates over a chain of three images (A-B, B-C, A-C): what find_matches accumulates
    a, b = make_pair(31, (0.0, 0.0, -95.0), 2.0, (20.0, 2.0, -95.0), 5.0, n=400)
    b2, c = make_pair(32, (20.0, 2.0, -95.0), 5.0, (41.0, 3.0, -94.0), 3.0, n=400)
    with contextlib.redirect_stdout(io.StringIO()):
        r1 = smart.update_surface_estimate(a, b)
        y1 = smart.update_yaw_error_estimate(a, b)
        y2 = smart.update_yaw_error_estimate(b, a)
        b2.name = b.name                      # same image B seen in the second pair
        c.match_list = {b.name: c.match_list.pop(list(c.match_list)[0])}
        b2.match_list = {c.name: b2.match_list.pop(list(b2.match_list)[0])}
        r2 = smart.update_surface_estimate(b2, c)
        y3 = smart.update_yaw_error_estimate(b2, c)
        ground_ab = smart.get_surface_estimate(a, b)
        ground_bc = smart.get_surface_estimate(b2, c)
    for tag, im in (("a", a), ("b", b), ("b2", b2), ("c", c)):
        out["chain_pts_" + tag] = np.float32([k.pt for k in im.kp_list])
        out["chain_ned_" + tag] = np.float64(im.ned)
        out["chain_yaw_" + tag] = im.yaw_deg
        out["chain_proj_" + tag] = np.concatenate([cv2.Rodrigues(im.get_proj()[0])[0], np.asarray(im.get_proj()[1])], 1)
    out["chain_matches_ab"] = np.int32(a.match_list[b.name])
    out["chain_matches_bc"] = np.int32(b2.match_list[c.name])
    out["chain_results"] = np.float64([r1[0], r1[1], y1, y2, r2[0], r2[1], y3, ground_ab, ground_bc,
                                       smart.get_yaw_error_estimate(a), smart.get_yaw_error_estimate(b)])
    out["yaw_bias"] = YAW_BIAS
    print("chain:", out["chain_results"])
    np.savez_compressed(os.path.join(HERE, "smart_reference.npz"), **out)


if __name__ == "__main__":
    main()
