"""GMS grid filter: the oracle restatement against the reference's own
scripts/lib/archive/gms_matcher.py (goldens from tests/golden/make_golden_gms.py),
and the CUDA kernel against both."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import oracle


def _case(g, name):
    return g[name + "_pts1"], g[name + "_pts2"], g[name + "_matches"], tuple(int(v) for v in g[name + "_size"])


def test_oracle_gms_equals_reference_module():
    g = load_golden("gms_reference.npz")
    for name in g["names"]:
        pts1, pts2, matches, size = _case(g, str(name))
        mask = oracle.gms_mask(pts1, pts2, size, size, matches, with_rotation=True, with_scale=False, threshold_factor=5.0)
        assert (mask == g[str(name) + "_mask"]).all(), name
    pts1, pts2, matches, size = _case(g, "flags")
    for ws, wr in ((False, False), (True, False), (True, True)):
        mask = oracle.gms_mask(pts1, pts2, size, size, matches, with_rotation=wr, with_scale=ws, threshold_factor=5.0)
        assert (mask == g["flags_s%d_r%d_mask" % (ws, wr)]).all(), (ws, wr)


def test_gms_last_half_cell_documents_both_rules():
    """Key points in the last half cell have no cell in the shifted grids (index -1).  The reference's archive Python
    (gms_matcher.py:205) reads mCellPairs[-1] there -- a wrap-around to cell 399; OpenCV's C++ matchGMS, which is what
    the call site (matcher.py:285) executes, skips such a match for that grid.  The oracle restates both:
    archive_wrap=True equals the unmodified archive module on the "edge_strip" fixture (a scene built so that the
    wrap-around really admits edge-strip matches), the default (C++ rule, what the CUDA kernel does) keeps a subset --
    it can only drop matches of that strip, never add one, and never touches a match elsewhere."""
    g = load_golden("gms_reference.npz")
    pts1, pts2, matches, size = _case(g, "edge_strip")
    ref = g["edge_strip_mask"]
    py = oracle.gms_mask(pts1, pts2, size, size, matches, archive_wrap=True)
    cxx = oracle.gms_mask(pts1, pts2, size, size, matches, archive_wrap=False)
    assert (py == ref).all()                              # the literal restatement is pinned by the module itself
    diff = np.nonzero(py != cxx)[0]
    assert len(diff) > 0                                  # the fixture does exercise the deviation
    assert (py[diff] & ~cxx[diff]).all()                  # the C++ rule only ever drops
    x = pts1[matches[diff, 0], 0] / size[0]
    y = pts1[matches[diff, 0], 1] / size[1]
    assert ((x >= 0.975) | (y >= 0.975)).all()            # ... and only matches whose key point lies in the last half cell
    # away from the strip (every other fixture) the two rules coincide
    for name in g["names"]:
        p1, p2, m, sz = _case(g, str(name))
        assert (oracle.gms_mask(p1, p2, sz, sz, m, archive_wrap=True) == g[str(name) + "_mask"]).all(), name


def test_oracle_pipeline_with_gms_equals_reference_module():
    """basic_pair_matches / bidirectional_pair_matches of the unmodified reference module with its GMS stage live
    (goldens: reference_gms_pipeline.npz) against the oracle's restated pipeline."""
    g = load_golden("reference_gms_pipeline.npz")
    size = (5472, 3648)
    kw = dict(norm=oracle.NORM_L2, match_ratio=0.75, max_distance=270.0, threads=4, size=size, dedupe=True)
    f = oracle.basic_pair(g["des0"], g["des1"], pts_q=g["pts0"], pts_t=g["pts1"], **kw)
    r = oracle.basic_pair(g["des1"], g["des0"], pts_q=g["pts1"], pts_t=g["pts0"], **kw)
    assert f == g["basic01"].tolist() and r == g["basic10"].tolist()
    b1, b2 = oracle.bidirectional(g["des1"], g["des2"], pts_q=g["pts1"], pts_t=g["pts2"], **kw)
    assert b1 == g["bidir12_fwd"].tolist() and b2 == g["bidir12_rev"].tolist()
    assert len(f) > 200 and len(b1) > 200


# ------------------------------------------------------------------ GPU
def _engine():
    from imageanalysis_b200 import _capi
    return _capi, _capi.Engine(_capi.NORM_L2, 128, 0)


@pytest.mark.gpu
def test_gpu_gms_equals_reference_module():
    _capi, eng = _engine()
    g = load_golden("gms_reference.npz")
    for name in g["names"]:
        pts1, pts2, matches, size = _case(g, str(name))
        mask = eng.gms_filter(pts1, pts2, matches, size)
        assert (mask == g[str(name) + "_mask"]).all(), name
    pts1, pts2, matches, size = _case(g, "flags")
    for ws, wr in ((False, False), (True, False), (True, True)):
        mask = eng.gms_filter(pts1, pts2, matches, size, with_rotation=wr, with_scale=ws)
        assert (mask == g["flags_s%d_r%d_mask" % (ws, wr)]).all(), (ws, wr)
    # the last-half-cell strip: the kernel follows OpenCV's C++ rule (skip), not the archive Python's wrap-around
    pts1, pts2, matches, size = _case(g, "edge_strip")
    mask = eng.gms_filter(pts1, pts2, matches, size)
    assert (mask == oracle.gms_mask(pts1, pts2, size, size, matches, archive_wrap=False)).all()
    assert (g["edge_strip_mask"] | ~mask).all() and (mask != g["edge_strip_mask"]).sum() > 0   # a strict subset of the archive result
    # ... and reproduces the archive module bit for bit when its rule is selected
    assert (eng.gms_filter(pts1, pts2, matches, size, archive_wrap=True) == g["edge_strip_mask"]).all()
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n,seed", [(1, 1), (30, 2), (700, 3), (2000, 4), (4096, 5)])
def test_gpu_gms_equals_oracle_random_scenes(n, seed):
    """Random scenes including points in the last half cell (no cell in the shifted grids), every match in one
    cell pair (hash contention, large counts), threshold factors and both flag settings."""
    _capi, eng = _engine()
    rng = np.random.default_rng(seed)
    size = (1920, 1080)
    p1 = rng.uniform(0, 1, (n, 2))
    ang = np.deg2rad(rng.choice([0, 90, 135, 180, 270]))
    R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
    p2 = np.clip((p1 - 0.5) @ R.T * 0.9 + 0.5 + rng.normal(0, 0.003, (n, 2)), 0, 0.9999)
    bad = rng.random(n) < 0.4
    p2[bad] = rng.uniform(0, 0.9999, (int(bad.sum()), 2))
    if n >= 700:
        p1[:300] = [0.31, 0.42] + rng.normal(0, 0.004, (300, 2))    # one crowded cell
        p2[:300] = [0.64, 0.22] + rng.normal(0, 0.004, (300, 2))
        p1[300:340, 0] = rng.uniform(0.976, 0.9999, 40)             # last half cell
    pts1 = (np.clip(p1, 0, 0.9999) * size).astype(np.float32)
    pts2 = (p2 * size).astype(np.float32)
    m = np.stack([rng.permutation(n), np.arange(n)], 1).astype(np.int32)
    pts1 = pts1[np.argsort(m[:, 0])]                                 # match k pairs pts1[m[k,0]] with pts2[k]
    for wr, ws, thr in ((True, False, 5.0), (False, False, 6.0), (True, True, 4.0)):
        want = oracle.gms_mask(pts1, pts2, size, size, m, with_rotation=wr, with_scale=ws, threshold_factor=thr)
        got = eng.gms_filter(pts1, pts2, m, size, with_rotation=wr, with_scale=ws, threshold_factor=thr)
        assert (got == want).all(), (wr, ws, thr, int(got.sum()), int(want.sum()))
    eng.close()


@pytest.mark.gpu
def test_gpu_match_pipeline_with_gms_equals_reference_module():
    """Device pipeline kNN -> metric -> GMS -> filter_duplicates -> gates -> cross-check against what the
    reference's own module produced with its GMS stage live."""
    import types
    from imageanalysis_b200 import matcher
    _capi, eng = _engine()
    g = load_golden("reference_gms_pipeline.npz")
    n = int(g["n"])
    for i in range(n):
        eng.upload(i, g["des%d" % i])
        eng.upload_keypoints(i, g["pts%d" % i])
        eng.upload_keypoint_keys(i, matcher.keypoint_keys([types.SimpleNamespace(pt=(float(x), float(y))) for x, y in g["pts%d" % i]]))
    pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]
    prm = _capi.Engine.make_params(dedupe=True, cross_check=True, gms=True, size=(5472, 3648))
    table, count = eng.match_pairs(pairs, prm)
    for p, (i, j) in enumerate(pairs):
        want = g["match_frame%02d_frame%02d" % (i, j)].tolist()
        assert table[p, :count[p]].tolist() == want, (i, j)
    prm = _capi.Engine.make_params(dedupe=True, cross_check=False, gms=True, size=(5472, 3648))
    table, count, rtable, rcount = eng.match_pairs([(0, 1)], prm, want_reverse=True)
    assert table[0, :count[0]].tolist() == g["basic01"].tolist()
    assert rtable[0, :rcount[0]].tolist() == g["basic10"].tolist()
    # one-call form: keypoints first, descriptors inside the call
    eng2 = _capi.Engine(_capi.NORM_L2, 128, 0)
    for i in range(n):
        eng2.upload_keypoints(i, g["pts%d" % i])
    prm = _capi.Engine.make_params(dedupe=False, cross_check=True, gms=True, size=(5472, 3648))
    t2, c2 = eng2.match_images(list(range(n)), [g["des%d" % i] for i in range(n)], pairs, prm)
    for p, (i, j) in enumerate(pairs):
        f, _ = oracle.bidirectional(g["des%d" % i], g["des%d" % j], oracle.NORM_L2, 0.75, 270.0, threads=4,
                                    pts_q=g["pts%d" % i], pts_t=g["pts%d" % j], size=(5472, 3648))
        assert t2[p, :c2[p]].tolist() == f, (i, j)
    eng.close()
    eng2.close()
