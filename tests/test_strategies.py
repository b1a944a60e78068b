"""The alternative matching strategies (matcher.py:358-850) and the robust-fit siblings of filter_by_transform
(matcher.py:121-126) against goldens produced by the UNMODIFIED reference module / live cv2
(tests/golden/make_golden_strategies.py).

RANSAC samplers differ between OpenCV and the CUDA kernel, so results are compared as SETS with a stated
tolerance: intersection-over-union >= 0.95 of the [queryIdx, trainIdx] pair sets (strategies) or of the inlier
masks (robust fits)."""
import types

import numpy as np
import pytest

from conftest import load_golden
from oracle import oracle

IOU_TOL = 0.95


def _iou(a, b):
    a, b = set(map(tuple, np.asarray(a).reshape(-1, 2).tolist())), set(map(tuple, np.asarray(b).reshape(-1, 2).tolist()))
    return len(a & b) / max(1, len(a | b))


def _configure(g):
    from imageanalysis_b200 import matcher
    from imageanalysis_b200.propshim import getNode
    det = getNode("/config/detector", True)
    det.setString("detector", "SIFT")
    det.setFloat("scale", 1.0)
    mn = getNode("/config/matcher", True)
    mn.setFloat("match_ratio", 0.75)
    mn.setFloat("min_pairs", 25)
    mn.setFloat("ground_m", 0.0)
    cam = getNode("/config/camera", True)
    cam.setInt("width_px", int(g["size"][0]))
    cam.setInt("height_px", int(g["size"][1]))
    cam.setLen("K", 9, 0.0)
    for i, v in enumerate(np.asarray(g["K"]).ravel()):
        cam.setFloatEnum("K", i, float(v))
    cam.setLen("dist_coeffs", 5, 0.0)
    return matcher


class PoseImage:
    """Duck-typed lib.image.Image (image.py:25-97, :507-553) built from a golden scene."""
    CAM2BODY = np.array([[0.0, 0.0, 1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])

    def __init__(self, name, g, s, side):
        from imageanalysis_b200 import matcher
        self.name = name
        self.des_list = g["s%d_des_%s" % (s, side)].astype(np.float32)
        pts, size, ang = g["s%d_pts_%s" % (s, side)], g["s%d_size_%s" % (s, side)], g["s%d_ang_%s" % (s, side)]
        # cv2.KeyPoint stores size and angle as float32
        self.kp_list = [types.SimpleNamespace(pt=(float(p[0]), float(p[1])), size=float(np.float32(z)), angle=float(np.float32(a)))
                        for p, z, a in zip(pts, size, ang)]
        self.match_list = {}
        self.ned = g["s%d_ned_%s" % (s, side)].tolist()
        self.quat = g["s%d_quat_%s" % (s, side)]
        self._R = matcher._quat_matrix(self.quat)

    def get_camera_pose(self, opt=False):
        return self.ned, [0.0, 0.0, 0.0], list(self.quat)

    def get_body2ned(self, opt=False):
        return self._R

    def get_cam2body(self):
        return self.CAM2BODY


# ------------------------------------------------------------------ CPU
def test_neighbour_selection_equals_literal_restatement():
    """The vectorised per-query selection of smart / bruteforce against the literal Python loops."""
    from imageanalysis_b200 import matcher
    rng = np.random.default_rng(0)
    n, m, k = 400, 350, 3
    idx = np.stack([rng.permutation(m)[:k] for _ in range(n)]).astype(np.int32)
    dist = np.sort(rng.uniform(50, 400, (n, k)).astype(np.float32), axis=1)
    dist[:40, 1] = dist[:40, 0]                      # ratio exactly 1
    dist[40:60] = np.float32([120, 130, 320])        # third neighbour beyond the distance limit
    idx[60:70, 2] = -1                               # a short list
    pts1 = rng.uniform(0, 4000, (n, 2)).astype(np.float32)
    pts2 = rng.uniform(0, 4000, (m, 2)).astype(np.float32)
    size1, size2 = rng.uniform(2, 9, n), rng.uniform(2, 9, m)
    size2[idx[:200, 0]] = size1[:200] * rng.uniform(0.85, 1.2, 200)
    ang = np.zeros(n)
    pred = (pts1 + rng.normal(0, 30, (n, 2))).astype(np.float32)
    for limit, p in ((290.0, None), (300.0, pred)):
        rows, bj, bd, va = matcher._best_of_neighbours(idx, dist, (pts1, size1, ang), (pts2, size2, np.zeros(m)), 0.75, limit, pred1=p)
        want = oracle.best_of_neighbours_literal(idx, dist, pts1, size1, pts2, size2, 0.75, limit, pred1=p)
        assert rows.tolist() == [w[0] for w in want] and bj.tolist() == [w[1] for w in want]
        assert np.allclose(bd, [w[2] for w in want], rtol=1e-6) and np.allclose(va, [w[3] for w in want], atol=1e-6)
        assert len(want) > 100


def test_pose_prediction_equals_reference_preliminary_homography():
    """smart_pair_matches' starting homography (matcher.py:359-454) from the poses alone, against the matrix the
    reference's own cv2.findHomography(..., 0) call returned: the 81 grid points map within 0.05 px."""
    g = load_golden("reference_strategies.npz")
    matcher = _configure(g)
    for s in (0, 1):
        a, b = PoseImage("A", g, s, "a"), PoseImage("B", g, s, "b")
        H = matcher._pose_prediction(a, b, est_rotation=False)
        H0 = g["s%d_H0" % s]
        w, h = int(g["size"][0]), int(g["size"][1])
        grid = np.array([[u, v, 1.0] for v in np.linspace(0, h, 9) for u in np.linspace(0, w, 9)])
        pa, pb = grid @ H.T, grid @ H0.T
        err = np.abs(pa[:, :2] / pa[:, 2:] - pb[:, :2] / pb[:, 2:]).max()
        assert err < 0.05, err


def test_homography_lsq_recovers_a_known_matrix():
    from imageanalysis_b200 import matcher
    rng = np.random.default_rng(1)
    H = np.array([[1.02, 0.03, 40.0], [-0.02, 0.98, -25.0], [1e-6, -2e-6, 1.0]])
    src = rng.uniform(0, 5000, (81, 2))
    d = np.concatenate([src, np.ones((81, 1))], 1) @ H.T
    got = matcher._fit_homography_lsq(src, d[:, :2] / d[:, 2:])
    assert np.allclose(got, H, rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("strategy", ["smart", "ratio", "bruteforce"])
def test_gpu_strategy_equals_reference_module(strategy):
    g = load_golden("reference_strategies.npz")
    matcher = _configure(g)
    matcher.configure()
    for s in (0, 1):
        a, b = PoseImage("A", g, s, "a"), PoseImage("B", g, s, "b")
        if strategy == "smart":
            fwd, rev = matcher.smart_pair_matches(a, b, review=False, est_rotation=False)
        elif strategy == "ratio":
            fwd, rev = matcher.ratio_pair_matches(a, b, review=False, est_rotation=False)
        else:
            fwd, rev = matcher.bruteforce_pair_matches(a, b, review=False)
        want = g["s%d_%s" % (s, strategy)]
        assert rev == [[t, q] for q, t in fwd]
        assert _iou(fwd, want) >= IOU_TOL, (strategy, s, len(fwd), len(want), _iou(fwd, want))
        # order convention of the reference: ascending position in the winning bin, i.e. query order for these strategies
        assert len(fwd) >= 25


@pytest.mark.gpu
def test_gpu_smart_without_poses_starts_from_identity():
    """Images without a pose API (the duck-typed minimum) must not raise: the prediction falls back to the identity
    (the default strategy of find_matches must always run, matcher.py:852)."""
    g = load_golden("reference_strategies.npz")
    matcher = _configure(g)
    matcher.configure()
    a, b = PoseImage("A", g, 0, "a"), PoseImage("B", g, 0, "b")
    for im in (a, b):
        im.get_body2ned = None
        del im.get_body2ned
    bare = [types.SimpleNamespace(name=im.name, des_list=im.des_list, kp_list=im.kp_list, match_list={}) for im in (a, b)]
    fwd, rev = matcher.smart_pair_matches(bare[0], bare[1])
    assert rev == [[t, q] for q, t in fwd]
    assert _iou(fwd, g["s0_smart"]) >= 0.9      # the 18 m baseline moves points by < 2048 px: the widest bin still finds the scene


@pytest.mark.gpu
@pytest.mark.parametrize("transform", ["homography", "fundamental"])
def test_gpu_filter_by_transform_equals_cv2(transform):
    """filter_by_transform (matcher.py:90-142) with cv2.findHomography / cv2.findFundamentalMat replaced by the GPU
    RANSAC: surviving matches against the masks live cv2 produced on the same points (IoU >= 0.95)."""
    g = load_golden("robust_fits.npz")
    gs = load_golden("reference_strategies.npz")
    matcher = _configure(gs)
    matcher.configure()
    key = "h" if transform == "homography" else "f"
    for s in range(3):
        p1, p2, mask = g["%s_p1_%d" % (key, s)], g["%s_p2_%d" % (key, s)], g["%s_mask_%d" % (key, s)].astype(bool)
        n = len(p1)
        i1 = types.SimpleNamespace(name="a", uv_list=p1.tolist(), match_list={"b": [[i, i] for i in range(n)]}, width=5472)
        i2 = types.SimpleNamespace(name="b", uv_list=p2.tolist(), match_list={})
        clean = matcher.filter_by_transform(g["K"], i1, i2, transform)
        kept = np.zeros(n, bool)
        kept[[m[0] for m in i1.match_list["b"]]] = True
        iou = (kept & mask).sum() / max(1, (kept | mask).sum())
        assert iou >= IOU_TOL, (transform, s, int(kept.sum()), int(mask.sum()), iou)
        assert clean == bool(kept.all())
        truth = g["%s_truth_%d" % (key, s)].astype(bool)
        assert kept[truth].mean() >= 0.97          # planted inliers survive
