"""CPU tests of the RANSAC minimal solvers (the same source the device code
compiles, run on the host through iam_debug_minimal_solver).  No GPU needed."""
import numpy as np

from conftest import load_golden
from imageanalysis_b200 import _capi, synth

K = np.array([[3666.666504, 0, 2736], [0, 3666.666504, 1824], [0, 0, 1]])


def _norm(p):
    return (p - [K[0, 2], K[1, 2]]) / [K[0, 0], K[1, 1]]


def test_five_point_recovers_the_true_essential_matrix():
    hits = 0
    for seed in range(12):
        p1, p2, _ = synth.two_view_scene(60, 0.0, K, seed=seed, noise_px=0.0)
        n1, n2 = _norm(p1), _norm(p2)
        Es = _capi.minimal_solver_host(_capi.MODEL_ESSENTIAL, n1[:5, 0], n1[:5, 1], n2[:5, 0], n2[:5, 1])
        assert 1 <= len(Es) <= 10
        h1, h2 = np.c_[n1, np.ones(60)], np.c_[n2, np.ones(60)]
        best = np.inf
        for E in Es:
            E = E.astype(np.float64)
            r = np.abs(np.einsum("ni,ij,nj->n", h2, E, h1))
            assert r[:5].max() < 1e-5                      # every root satisfies the 5 sample constraints
            assert abs(np.linalg.det(E)) < 1e-6            # and the cubic constraints
            assert np.abs(2 * E @ E.T @ E - np.trace(E @ E.T) * E).max() < 1e-5
            best = min(best, np.median(r))
        hits += best < 1e-6                                # one root is the true geometry (all 60 points fit)
    assert hits >= 11


def test_five_point_agrees_with_cv2_essential_on_golden_inliers():
    """On the planted inliers of the golden scene (cv2.findEssentialMat output
    recorded by make_golden.py) the solver's best root explains the same points."""
    g = load_golden("find_essential.npz")
    p1, p2, mask = g["p1_0"], g["p2_0"], g["mask_0"].astype(bool)
    n1, n2 = _norm(p1[mask]), _norm(p2[mask])
    thr = (float(g["tol"]) / K[0, 0]) ** 2
    h1, h2 = np.c_[n1, np.ones(len(n1))], np.c_[n2, np.ones(len(n2))]
    best = 0
    for s in range(0, 50, 5):
        Es = _capi.minimal_solver_host(0, n1[s:s + 5, 0], n1[s:s + 5, 1], n2[s:s + 5, 0], n2[s:s + 5, 1])
        for E in Es.astype(np.float64):
            Ex1 = h1 @ E.T
            Etx2 = h2 @ E
            r = np.einsum("ni,ni->n", h2, Ex1)
            samp = r * r / (Ex1[:, 0] ** 2 + Ex1[:, 1] ** 2 + Etx2[:, 0] ** 2 + Etx2[:, 1] ** 2)
            best = max(best, (samp <= thr).mean())
    assert best > 0.95


def test_four_point_homography_exact():
    rng = np.random.default_rng(0)
    for _ in range(10):
        H = np.eye(3) + rng.normal(0, 0.05, (3, 3))
        H /= H[2, 2]
        a = rng.uniform(-1, 1, (4, 2))
        b = (H @ np.c_[a, np.ones(4)].T).T
        b = b[:, :2] / b[:, 2:]
        got = _capi.minimal_solver_host(_capi.MODEL_HOMOGRAPHY, a[:, 0], a[:, 1], b[:, 0], b[:, 1])
        assert len(got) == 1 and np.abs(got[0] - H).max() < 1e-3


def test_degenerate_samples_return_no_model():
    z = np.zeros(5, np.float32)
    assert len(_capi.minimal_solver_host(0, z, z, z, z)) == 0
    assert len(_capi.minimal_solver_host(1, z, z, z, z)) == 0


def test_seven_point_recovers_the_true_fundamental_matrix():
    """The 7-point solver behind the 'fundamental' branch (matcher.py:124): every root is rank 2 and satisfies the
    seven sample constraints; one of them is the true geometry (all 60 noise-free points fit)."""
    hits = 0
    for seed in range(12):
        p1, p2, _ = synth.two_view_scene(60, 0.0, K, seed=seed, noise_px=0.0)
        n1, n2 = _norm(p1), _norm(p2)          # any well-conditioned coordinates do: F is not tied to K
        Fs = _capi.minimal_solver_host(_capi.MODEL_FUNDAMENTAL, n1[:7, 0], n1[:7, 1], n2[:7, 0], n2[:7, 1])
        assert 1 <= len(Fs) <= 3
        h1, h2 = np.c_[n1, np.ones(60)], np.c_[n2, np.ones(60)]
        best = np.inf
        for F in Fs.astype(np.float64):
            r = np.abs(np.einsum("ni,ij,nj->n", h2, F, h1))
            assert r[:7].max() < 1e-5
            assert abs(np.linalg.det(F)) < 1e-7
            best = min(best, np.median(r))
        hits += best < 1e-5
    assert hits >= 11
    z = np.zeros(7, np.float32)
    assert len(_capi.minimal_solver_host(_capi.MODEL_FUNDAMENTAL, z, z, z, z)) == 0
