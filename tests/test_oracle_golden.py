"""CPU tests: the oracle against the committed golden vectors (live cv2 output
and the reference's own lib.matcher run with shims; tests/golden/make_golden.py),
the host logic, and the C-ABI library's exports.  No GPU needed."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden
from oracle import oracle


@pytest.fixture(scope="module", autouse=True)
def _build_oracle():
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    assert oracle.have_c_oracle()


KNN_CASES = [("knn_l2_synth.npz", oracle.NORM_L2), ("knn_hamming_synth.npz", oracle.NORM_HAMMING),
             ("knn_sift_real.npz", oracle.NORM_L2), ("knn_orb_real.npz", oracle.NORM_HAMMING)]


@pytest.mark.parametrize("name,norm", KNN_CASES)
@pytest.mark.parametrize("k", [1, 2, 3])
def test_c_oracle_equals_cv2_bfmatcher(name, norm, k):
    g = load_golden(name)
    idx, dist = oracle.knn(g["q"], g["t"], k, norm, threads=2)
    assert (idx == g["idx"][:, :k]).all()
    assert (dist == g["dist"][:, :k]).all()          # bit-exact float32 distances
    ridx, rdist = oracle.knn(g["t"], g["q"], k, norm, threads=3)
    assert (ridx == g["ridx"][:, :k]).all()
    assert (rdist == g["rdist"][:, :k]).all()


@pytest.mark.parametrize("name,norm", KNN_CASES[:2])
def test_numpy_oracle_equals_c_oracle(name, norm):
    g = load_golden(name)
    a = oracle.knn_numpy(g["q"][:150], g["t"], 3, norm)
    b = oracle.knn(g["q"][:150], g["t"], 3, norm)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()


def test_ties_resolve_to_lowest_train_index():
    g = load_golden("knn_l2_synth.npz")
    # q[302] == t[100] == t[200]: cv2 reports 100 then 200, both at distance 0
    assert g["idx"][302, 0] == 100 and g["idx"][302, 1] == 200
    assert g["dist"][302, 0] == 0 and g["dist"][302, 1] == 0
    idx, _ = oracle.knn(g["q"], g["t"], 2, oracle.NORM_L2)
    assert idx[302].tolist() == [100, 200]


def test_float_path_within_tolerance():
    rng = np.random.default_rng(0)
    q = rng.normal(size=(50, 64)).astype(np.float32)
    t = rng.normal(size=(70, 64)).astype(np.float32)
    idx, dist = oracle.knn(q, t, 2, oracle.NORM_L2)
    ref = np.sqrt(((q[:, None].astype(np.float64) - t[None].astype(np.float64)) ** 2).sum(-1))
    assert np.allclose(dist[:, 0], ref.min(1), rtol=1e-6)


def test_fewer_train_rows_than_k():
    q = np.zeros((3, 32), np.uint8)
    t = np.ones((1, 32), np.uint8)
    idx, dist = oracle.knn(q, t, 2, oracle.NORM_HAMMING)
    assert (idx[:, 0] == 0).all() and (idx[:, 1] == -1).all() and np.isinf(dist[:, 1]).all()


def test_metric_reduction_equals_reference_module():
    """matcher.py:253-269 as executed by the reference itself (GMS replaced by an
    identity that recorded its input)."""
    g = load_golden("reference_reductions.npz")
    for a, b, key in ((0, 1, "thresh01"), (1, 0, "thresh10")):
        idx, dist = oracle.knn(g["des%d" % a], g["des%d" % b], 2, oracle.NORM_L2, threads=2)
        got = oracle.reduce_ref_metric(idx, dist, 0.75, 270.0, 2000, 25)
        assert got == g[key].tolist()


def test_dedupe_and_cross_check_equal_reference_module():
    g = load_golden("reference_reductions.npz")
    d01 = oracle.filter_duplicates(g["pts0"], g["pts1"], g["thresh01"].tolist())
    d10 = oracle.filter_duplicates(g["pts1"], g["pts0"], g["thresh10"].tolist())
    assert d01 == g["basic01"].tolist() and d10 == g["basic10"].tolist()
    assert len(d01) < len(g["thresh01"])     # the fixture really contains duplicates
    c1, c2 = oracle.filter_cross_check(d01, d10)
    assert c1 == g["cross01"].tolist() and c2 == g["cross10"].tolist()


def test_min_pairs_gate_returns_empty():
    g = load_golden("reference_reductions.npz")
    f, r = oracle.bidirectional(g["des0"], g["des_far"], oracle.NORM_L2, 0.75, 270.0)
    assert f == [] and r == [] and g["bidir_far_fwd"].shape[0] == 0


def test_worklist_counts_match_survey():
    from imageanalysis_b200 import pairs, synth
    neds = synth.survey_grid_neds()
    assert len(neds) == 2812
    seq = pairs.worklist(neds, "sequential")
    assert len(seq) == 4 * 2812 - 10                    # 11 238
    small = synth.survey_grid_neds(5, 8)
    for mode in ("sequential", "geotag"):
        assert pairs.worklist(small, mode) == oracle.worklist(small, mode)


def test_shard_partition_is_exact():
    from imageanalysis_b200 import pairs
    for n in (0, 1, 7, 1990, 11238):
        for w in (1, 2, 3, 8):
            cuts = [pairs.shard(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_keypoint_keys_equal_string_keys():
    from imageanalysis_b200 import matcher
    import types
    rng = np.random.default_rng(3)
    pts = rng.uniform(0, 5000, (500, 2)).astype(np.float32)
    pts[100:120] = pts[:20]
    pts[200] = pts[0] + np.float32(0.004)        # same 2-decimal key unless it crosses a rounding boundary
    kps = [types.SimpleNamespace(pt=(float(x), float(y))) for x, y in pts]
    keys = matcher.keypoint_keys(kps)
    strs = ["%.2f-%.2f" % k.pt for k in kps]
    for i in range(0, 500, 7):
        for j in range(500):
            assert (keys[i] == keys[j]) == (strs[i] == strs[j])


def test_propshim_roundtrip(tmp_path):
    from imageanalysis_b200 import propshim
    n = propshim.getNode("/unit/test", True)
    n.setFloat("scale", 0.4)
    n.setString("detector", "SIFT")
    n.setLen("K", 9, 0.0)
    n.setFloatEnum("K", 4, 3666.5)
    assert propshim.getNode("/unit/test").getFloat("scale") == 0.4
    assert n.getFloatEnum("K", 4) == 3666.5 and n.getLen("K") == 9
    assert n.getString("missing") == "" and n.getFloat("missing") == 0.0
    f = tmp_path / "c.json"
    assert propshim.save(str(f), propshim.getNode("/unit"))
    m = propshim.PropertyNode()
    assert propshim.load(str(f), m) and m.getChild("test").getString("detector") == "SIFT"


def test_library_exports_every_declared_symbol():
    """include/iamatch.h <-> libiamatch.so <-> _capi.EXPORTS agree; loading needs no GPU."""
    from imageanalysis_b200 import _capi, build
    build.build(verbose=False)
    lib = ctypes.CDLL(_capi.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "iamatch.h")).read()
    declared = set(re.findall(r"\b(iam_[a-z_]+)\s*\(", header))
    assert declared == set(_capi.EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    lib.iam_abi_version.restype = ctypes.c_int
    assert lib.iam_abi_version() == 6


def test_product_fails_loudly_without_gpu_or_library(monkeypatch):
    from imageanalysis_b200 import _capi
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_capi.IamError):
        _capi.Engine(_capi.NORM_L2, 128, 0)
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", "/nonexistent/libiamatch.so")
    with pytest.raises(_capi.IamError, match="no CPU fallback"):
        _capi.load_library()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "imageanalysis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("checker", ""), f
