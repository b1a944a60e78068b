"""Pair-wise side estimators (reference scripts/lib/smart.py): triangulation of a pair's matches + surface statistics
(:26-63, :116-131), the partial-affine fit + yaw error (:66-90, :139-190), and the running property-tree estimates
(:194-317), against goldens recorded by running the UNMODIFIED reference module (tests/golden/make_golden_smart.py).

Tolerances: triangulated points 1e-6 m against cv2.triangulatePoints on well-conditioned points (float64 Jacobi SVD
here, OpenCV's SVD there; points whose rays are nearly parallel -- planted outliers -- are excluded by a depth
window), surface mean / std 1e-6 m on scenes without outliers.  cv2.estimateAffinePartial2D's RANSAC sampler cannot
be reproduced: rotation within 0.05 degree, scale within 2e-3, yaw error within 0.1 degree, inlier sets IoU >= 0.8 on
the flat scene (on rolling terrain a similarity explains only part of the matches and RANSAC may settle on another
consistent subset: there only the decomposed parameters are compared)."""
import numpy as np
import pytest

from conftest import load_golden


class PoseImage:
    """Duck-typed lib.image.Image: what smart.py touches."""

    def __init__(self, name, pts, ned, proj=None, yaw=0.0, bias=0.0):
        from imageanalysis_b200 import detector
        n = len(pts)
        self.name = name
        self.kp_list = detector._keypoints(dict(pt=pts, size=np.full(n, 3.0), angle=np.zeros(n), response=np.zeros(n),
                                                octave=np.zeros(n, np.int32)))
        self.match_list = {}
        self.ned = [float(v) for v in ned]
        self.proj = proj
        self.yaw, self.bias = float(yaw), float(bias)
        self.yaw_error = None

    def get_camera_pose(self, opt=False):
        return self.ned, [0.0, 0.0, 0.0], [1.0, 0.0, 0.0, 0.0]

    def get_aircraft_pose(self):
        return [0.0, 0.0, 0.0], [self.yaw - self.bias, 0.0, 0.0], [1.0, 0.0, 0.0, 0.0]

    def set_aircraft_yaw_error_estimate(self, v):
        self.yaw_error = v

    def get_proj(self, opt=False, yaw_error_est=0.0):
        R, t = self.proj[:, :3], self.proj[:, 3:]
        return _rodrigues_vector(R), t


def _rodrigues_vector(R):
    th = np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1))
    if th < 1e-12:
        return np.zeros((3, 1))
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (2 * np.sin(th))
    return (w * th).reshape(3, 1)


def _configure(g):
    from imageanalysis_b200.propshim import getNode
    cam = getNode("/config/camera", True)
    cam.setInt("width_px", int(g["size"][0]))
    cam.setInt("height_px", int(g["size"][1]))
    cam.setLen("K", 9, 0.0)
    for i, v in enumerate(g["K"].ravel()):
        cam.setFloatEnum("K", i, float(v))


def _scene(g, s):
    pre = "s%d_" % s
    bias = float(g["yaw_bias"])
    a = PoseImage("A", g[pre + "pts_a"], g[pre + "ned_a"], g[pre + "proj_a"], float(g[pre + "yaw_a"]), bias)
    b = PoseImage("B", g[pre + "pts_b"], g[pre + "ned_b"], g[pre + "proj_b"], float(g[pre + "yaw_b"]), bias)
    m = g[pre + "matches"].tolist()
    a.match_list["B"] = m
    b.match_list["A"] = [[q[1], q[0]] for q in m]
    return a, b


def test_golden_is_consistent_with_its_own_scene():
    """numpy restatement of the linear triangulation on the golden inputs reproduces the recorded points (pins the
    input conventions: [R | t] rows, K^-1 normalisation, match order)."""
    g = load_golden("smart_reference.npz")
    IK = np.linalg.inv(g["K"])
    for s in range(3):
        pre = "s%d_" % s
        m = g[pre + "matches"]
        uv1 = np.concatenate([g[pre + "pts_a"][m[:, 0]].astype(np.float64), np.ones((len(m), 1))], 1)
        uv2 = np.concatenate([g[pre + "pts_b"][m[:, 1]].astype(np.float64), np.ones((len(m), 1))], 1)
        x1, x2 = (IK @ uv1.T)[:2].T, (IK @ uv2.T)[:2].T
        P1, P2 = g[pre + "proj_a"], g[pre + "proj_b"]
        want = g[pre + "points"]
        for i in range(0, len(m), 97):
            A = np.stack([x1[i, 0] * P1[2] - P1[0], x1[i, 1] * P1[2] - P1[1], x2[i, 0] * P2[2] - P2[0], x2[i, 1] * P2[2] - P2[1]])
            X = np.linalg.svd(A)[2][3]
            X = X / X[3]
            if abs(want[2, i]) < 500:
                assert np.allclose(X[:3], want[:3, i], rtol=0, atol=1e-6)
        assert np.allclose(want[3], 1.0)


def test_decompose_affine_and_tree_bookkeeping_without_gpu():
    from imageanalysis_b200 import smart
    rot, tx, ty, sx, sy = smart.decompose_affine(np.array([[0.8, -0.6, 5.0], [0.6, 0.8, -7.0]]))
    assert abs(rot - 36.86989764584402) < 1e-12 and (tx, ty) == (5.0, -7.0) and abs(sx - 1) < 1e-12 and abs(sy - 1) < 1e-12
    rot, _, _, sx, sy = smart.decompose_affine(np.array([[-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]]))
    assert sx == -1.0 and sy == -1.0 and abs(abs(rot) - 180.0) < 1e-9 or rot == 0.0
    a, b = PoseImage("ta", np.zeros((0, 2)), (0, 0, 0)), PoseImage("tb", np.zeros((0, 2)), (0, 0, 0))
    n1, n2 = smart.smart_node.getChild("ta", True), smart.smart_node.getChild("tb", True)
    n1.setFloat("srtm_surface_m", 10.0)
    n2.setFloat("srtm_surface_m", 20.0)
    assert smart.get_surface_estimate(a, b) == 15.0          # SRTM fall-back (smart.py:311-317)
    n1.setFloat("tri_surface_m", 31.0)
    assert smart.get_surface_estimate(a, b) == 31.0          # one triangulated side wins (:303-309)
    n2.setFloat("tri_surface_m", 33.0)
    assert smart.get_surface_estimate(a, b) == 32.0
    assert smart.get_yaw_error_estimate(a) == 0.0
    n1.setFloat("yaw_error", 2.5)
    assert smart.get_yaw_error_estimate(a) == 2.5
    assert smart.triangulate_features(a, a) is None and smart.find_affine(a, b) is None     # no matches recorded


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("s", [0, 1, 2])
def test_gpu_triangulation_equals_reference(s):
    from imageanalysis_b200 import smart
    g = load_golden("smart_reference.npz")
    _configure(g)
    a, b = _scene(g, s)
    pts = smart.triangulate_features(a, b)
    want = g["s%d_points" % s]
    assert pts.shape == want.shape and np.all(pts[3] == 1.0)
    sane = (np.abs(want[2]) < 500) & (np.abs(want[:3]).max(axis=0) < 5000)     # planted outliers triangulate anywhere
    assert sane.sum() >= 0.75 * want.shape[1]
    assert np.abs(pts[:3, sane] - want[:3, sane]).max() < 1e-6
    surf, std, dist = smart.estimate_surface_elevation(a, b)
    ws, wstd, wdist = g["s%d_surface" % s]
    assert abs(dist - wdist) < 1e-12
    if s == 0:
        assert abs(surf - ws) < 1e-6 and abs(std - wstd) < 1e-6
    else:      # ill-conditioned outlier points dominate the statistics: compare on the scale of the spread
        assert abs(surf - ws) < 1e-6 * max(1.0, wstd * 100) and abs(std - wstd) < 1e-6 * max(1.0, wstd * 100)


@pytest.mark.gpu
def test_gpu_surface_estimates_batched_equals_single():
    from imageanalysis_b200 import smart
    g = load_golden("smart_reference.npz")
    _configure(g)
    pairs = [_scene(g, s) for s in range(3)]
    empty_a, empty_b = PoseImage("EA", np.zeros((0, 2)), (0, 0, -90)), PoseImage("EB", np.zeros((0, 2)), (3, 4, -90))
    batch = smart.surface_estimates(pairs[:2] + [(empty_a, empty_b)] + pairs[2:])
    assert batch[2] == (None, None, 5.0)
    for one, (a, b) in zip(batch[:2] + batch[3:], pairs):
        assert one == smart.estimate_surface_elevation(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("s", [0, 1, 2])
def test_gpu_partial_affine_and_yaw_error(s):
    from imageanalysis_b200 import _capi, smart
    g = load_golden("smart_reference.npz")
    _configure(g)
    a, b = _scene(g, s)
    aff = smart.find_affine(a, b)
    want = g["s%d_affine" % s]
    assert aff.shape == (2, 3)
    rot, tx, ty, sx, sy = smart.decompose_affine(aff)
    wrot, wtx, wty, wsx, wsy = smart.decompose_affine(want)
    assert abs(rot - wrot) < 0.05 and abs(sx - wsx) < 2e-3 and abs(sy - wsy) < 2e-3
    assert abs(aff[0, 0] - aff[1, 1]) < 1e-6 and abs(aff[0, 1] + aff[1, 0]) < 1e-6          # a similarity
    yaw = smart.estimate_yaw_error(a, b)
    wyaw = g["s%d_yaw" % s]
    assert abs(yaw[0] - wyaw[0]) < 0.1 and abs(yaw[1] - wyaw[1]) < 1e-9 and abs(yaw[2] - wyaw[2]) < 0.1
    assert abs(yaw[0] - float(g["yaw_bias"])) < 0.1          # the planted EKF yaw bias is recovered
    yaw_rev = smart.estimate_yaw_error(b, a)
    assert abs(yaw_rev[0] - g["s%d_yaw_rev" % s][0]) < 0.1
    if s == 2:      # flat ground: the similarity explains every true match; compare the inlier sets
        m = g["s2_matches"]
        uv1, uv2 = g["s2_pts_a"][m[:, 0]], g["s2_pts_b"][m[:, 1]]
        mask, model, inl = smart._eng().ransac_pairs(_capi.MODEL_AFFINE_PARTIAL, uv2, uv1, np.array([0, len(m)], np.int32), None,
                                                     3.0, prob=0.99, max_iters=2000)
        wm = g["s2_affine_inliers"].astype(bool)
        iou = (mask.astype(bool) & wm).sum() / (mask.astype(bool) | wm).sum()
        assert iou >= 0.8 and abs(int(inl[0]) - int(wm.sum())) <= 0.1 * wm.sum()
        assert np.abs(model[0, :2, :2] - want[:, :2]).max() < 2e-3 and np.abs(model[0, :2, 2] - want[:, 2]).max() < 2.0


@pytest.mark.gpu
def test_gpu_running_estimates_over_a_chain():
    from imageanalysis_b200 import smart
    g = load_golden("smart_reference.npz")
    _configure(g)
    bias = float(g["yaw_bias"])

    def make(tag, name):
        return PoseImage(name, g["chain_pts_" + tag], g["chain_ned_" + tag], g["chain_proj_" + tag], float(g["chain_yaw_" + tag]), bias)

    a, b, b2, c = make("a", "CA"), make("b", "CB"), make("b2", "CB"), make("c", "CC")
    mab, mbc = g["chain_matches_ab"].tolist(), g["chain_matches_bc"].tolist()
    a.match_list["CB"] = mab
    b.match_list["CA"] = [[q[1], q[0]] for q in mab]
    b2.match_list["CC"] = mbc
    c.match_list["CB"] = [[q[1], q[0]] for q in mbc]
    want = g["chain_results"]
    r1 = smart.update_surface_estimate(a, b)
    y1 = smart.update_yaw_error_estimate(a, b)
    y2 = smart.update_yaw_error_estimate(b, a)
    r2 = smart.update_surface_estimate(b2, c)
    y3 = smart.update_yaw_error_estimate(b2, c)
    got = [r1[0], r1[1], y1, y2, r2[0], r2[1], y3, smart.get_surface_estimate(a, b), smart.get_surface_estimate(b2, c),
           smart.get_yaw_error_estimate(a), smart.get_yaw_error_estimate(b)]
    assert np.abs(np.array(got[:2]) - want[:2]).max() < 1e-6 and np.abs(np.array(got[4:6]) - want[4:6]).max() < 1e-6
    assert np.abs(np.array([got[2], got[3], got[6], got[9], got[10]]) - want[[2, 3, 6, 9, 10]]).max() <= 0.1001
    assert got[7] == want[7] and got[8] == want[8]            # averages of values rounded to 0.1 m
