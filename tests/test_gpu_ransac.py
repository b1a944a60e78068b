"""GPU tests of the batched RANSAC kernels against cv2.findEssentialMat output
recorded in tests/golden/find_essential.npz (generated exactly as the reference
calls it, matcher.py:126).  The sampler differs from OpenCV's, so parity is
defined on inlier SETS: IoU >= 0.95 and every clean planted inlier kept."""
import types

import numpy as np
import pytest

from conftest import load_golden
from imageanalysis_b200 import _capi, synth

pytestmark = pytest.mark.gpu


def _sampson(E, K, p1, p2):
    n1 = (p1 - [K[0, 2], K[1, 2]]) / [K[0, 0], K[1, 1]]
    n2 = (p2 - [K[0, 2], K[1, 2]]) / [K[0, 0], K[1, 1]]
    h1, h2 = np.c_[n1, np.ones(len(n1))], np.c_[n2, np.ones(len(n2))]
    Ex1, Etx2 = h1 @ E.T, h2 @ E
    r = np.einsum("ni,ni->n", h2, Ex1)
    return r * r / (Ex1[:, 0] ** 2 + Ex1[:, 1] ** 2 + Etx2[:, 0] ** 2 + Etx2[:, 1] ** 2)


def test_essential_inlier_sets_agree_with_cv2():
    g = load_golden("find_essential.npz")
    K, tol = g["K"], float(g["tol"])
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    p1 = np.concatenate([g["p1_%d" % s] for s in range(3)])
    p2 = np.concatenate([g["p2_%d" % s] for s in range(3)])
    off = np.cumsum([0] + [len(g["p1_%d" % s]) for s in range(3)]).astype(np.int32)
    mask, E, ninl = eng.ransac_pairs(_capi.MODEL_ESSENTIAL, p1, p2, off, K, tol)
    for s in range(3):
        m = mask[off[s]:off[s + 1]].astype(bool)
        ref = g["mask_%d" % s].astype(bool)
        truth = g["truth_%d" % s].astype(bool)
        iou = (m & ref).sum() / max(1, (m | ref).sum())
        assert iou >= 0.95, (s, iou)
        assert ninl[s] == m.sum()
        thr2 = (tol / ((K[0, 0] + K[1, 1]) / 2)) ** 2
        err = _sampson(E[s], K, g["p1_%d" % s], g["p2_%d" % s])
        assert ((err <= thr2) == m).mean() > 0.995          # the mask is the model's own inlier set
        clean = truth & (_sampson(g["E_%d" % s], K, g["p1_%d" % s], g["p2_%d" % s]) < 0.25 * thr2)
        assert m[clean].mean() > 0.98                        # clean planted inliers are kept
        # random outliers are rejected (a stray one may fit the geometry by chance, as it does for cv2)
        assert m[~truth].mean() < 0.15 or m[~truth].sum() <= ref[~truth].sum() + 1
    # deterministic call to call (fixed seed), like cv2.findEssentialMat
    mask2, E2, _ = eng.ransac_pairs(_capi.MODEL_ESSENTIAL, p1, p2, off, K, tol)
    assert (mask == mask2).all() and np.array_equal(E, E2)


def test_homography_ransac_planar_scene():
    rng = np.random.default_rng(4)
    H = np.array([[1.01, 0.02, 40.0], [-0.015, 0.99, -25.0], [2e-6, -1e-6, 1.0]])
    p1 = np.stack([rng.uniform(0, 5472, 800), rng.uniform(0, 3648, 800)], 1)
    q = (H @ np.c_[p1, np.ones(800)].T).T
    p2 = q[:, :2] / q[:, 2:] + rng.normal(0, 0.5, (800, 2))
    bad = rng.permutation(800)[:240]
    p2[bad] = np.stack([rng.uniform(0, 5472, 240), rng.uniform(0, 3648, 240)], 1)
    truth = np.ones(800, bool)
    truth[bad] = False
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    mask, Hg, ninl = eng.ransac_pairs(_capi.MODEL_HOMOGRAPHY, p1, p2, np.int32([0, 800]), None, 5.0)
    m = mask.astype(bool)
    assert m[truth].mean() > 0.97 and m[~truth].mean() < 0.05
    assert np.abs(Hg[0] / Hg[0][2, 2] - H).max() < 0.5 and abs(Hg[0][0, 0] - 1.01) < 0.01


def test_small_and_empty_sets():
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    K = np.array([[1000.0, 0, 500], [0, 1000.0, 400], [0, 0, 1]])
    p = np.random.default_rng(0).uniform(0, 1000, (3, 2)).astype(np.float32)
    mask, E, ninl = eng.ransac_pairs(_capi.MODEL_ESSENTIAL, p, p, np.int32([0, 3, 3]), K, 3.0)
    assert mask.sum() == 0 and (ninl == 0).all() and (E == 0).all()


def test_filter_by_transform_drops_outliers():
    from test_gpu_parity import FakeImage, _configure_matcher
    matcher = _configure_matcher()
    K = np.array([[3666.666504, 0, 2736], [0, 3666.666504, 1824], [0, 0, 1]])
    p1, p2, truth = synth.two_view_scene(600, 0.25, K, seed=42)
    i1 = FakeImage("a", None, p1, [0, 0, 0])
    i2 = FakeImage("b", None, p2, [15, 0, 0])
    i1.width = 5472
    i1.match_list["b"] = [[k, k] for k in range(600)]
    clean = matcher.filter_by_transform(K, i1, i2, "essential")
    kept = {q for q, t in i1.match_list["b"]}
    assert not clean
    assert len(kept & set(np.nonzero(truth)[0])) > 0.95 * truth.sum()
    assert len(kept & set(np.nonzero(truth == 0)[0])) < 0.15 * (600 - truth.sum())
    i1.match_list["b"] = [[k, k] for k in range(10)]       # below min_pairs: cleared (matcher.py:99-101)
    assert matcher.filter_by_transform(K, i1, i2, "essential") and i1.match_list["b"] == []


def test_ransac_on_device_tables_equals_the_host_point_form():
    """iam_ransac_tables: filter_by_transform for every pair of a match call without leaving the device -- the
    correspondences are the table rows looked up in the resident key points.  Same sampler, same pair numbering:
    masks and models must equal iam_ransac_pairs on the points gathered by hand; compact=True leaves exactly the
    inlier rows in the tables (order kept) and empties pairs below min_pairs (matcher.py:99-101, :134-141)."""
    K = np.array([[3666.666504, 0, 2736], [0, 3666.666504, 1824], [0, 0, 1]])
    tol = 5472 ** 0.25
    n = 1500
    des, pts, neds = synth.sift_project(3, n, seed=21, planted=0.4)
    rng = np.random.default_rng(5)
    # key points: random, except that matched rows of frames (0, 1) and (1, 2) follow a two-view geometry
    kp = [np.stack([rng.uniform(0, 5472, n), rng.uniform(0, 3648, n)], 1).astype(np.float32) for _ in range(3)]
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    for i in range(3):
        eng.upload(i, des[i])
    pairs = [(0, 1), (1, 2), (0, 2)]
    prm = _capi.Engine.make_params()
    table, count = eng.match_pairs(pairs, prm)
    for p, (a, b) in enumerate(pairs[:2]):
        c = int(count[p])
        p1, p2, _ = synth.two_view_scene(c, 0.25, K, seed=70 + p)
        kp[b][table[p, :c, 1]] = p2
        if p == 0:
            kp[a][table[p, :c, 0]] = p1
        else:       # frame 1's points are shared with pair 0: keep them, move frame 2 consistently instead
            q1 = kp[a][table[p, :c, 0]]
            kp[b][table[p, :c, 1]] = q1 + (p2 - p1)
    for i in range(3):
        eng.upload_keypoints(i, kp[i])
    eng.match_pairs_device(np.int32(pairs), prm)
    cap = prm.cap
    mask_t, E_t, inl_t = eng.ransac_tables(_capi.MODEL_ESSENTIAL, K, tol, len(pairs), cap, min_pairs=25, compact=False, want_mask=True)
    off = np.concatenate([[0], np.cumsum(count)]).astype(np.int32)
    g1 = np.concatenate([kp[a][table[p, :count[p], 0]] for p, (a, b) in enumerate(pairs)])
    g2 = np.concatenate([kp[b][table[p, :count[p], 1]] for p, (a, b) in enumerate(pairs)])
    mask_h, E_h, inl_h = eng.ransac_pairs(_capi.MODEL_ESSENTIAL, g1, g2, off, K, tol)
    for p in range(len(pairs)):
        c = int(count[p])
        if c < 25:
            assert inl_t[p] == 0
            continue
        assert (mask_t[p, :c] == mask_h[off[p]:off[p + 1]]).all(), p
        assert inl_t[p] == inl_h[p] and np.allclose(E_t[p], E_h[p], atol=1e-6)
    assert inl_t[0] > 0.6 * count[0]
    # in-place compaction
    eng.ransac_tables(_capi.MODEL_ESSENTIAL, K, tol, len(pairs), cap, min_pairs=25, compact=True, want_model=False)
    t2, c2 = eng.fetch_tables(len(pairs), cap)
    for p in range(len(pairs)):
        c = int(count[p])
        want = table[p, :c][mask_t[p, :c].astype(bool)] if c >= 25 else np.zeros((0, 2), np.int32)
        assert c2[p] == len(want) and (t2[p, :c2[p]] == want).all(), p
    eng.close()
