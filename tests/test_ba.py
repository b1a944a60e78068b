"""Bundle-adjustment residual / Jacobian: the oracle against the reference's own Optimizer.fun (goldens from
tests/golden/make_golden_ba.py) and the CUDA kernel against both."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import oracle


def _case(g, name):
    n_cam, n_pts = (int(v) for v in g[name + "_n"])
    K = g[name + "_K"]
    return dict(params=g[name + "_params"], n_cam=n_cam, n_pts=n_pts, cam_idx=g[name + "_cam_idx"],
                pt_idx=g[name + "_pt_idx"], obs_uv=g[name + "_uv"], K4=(K[0, 0], K[1, 1], K[0, 2], K[1, 2]),
                dist=g[name + "_dist"])


def test_oracle_residual_equals_reference_optimizer_fun():
    """float64 against cv2.projectPoints inside the unmodified Optimizer.fun: 1e-9 px absolute (the reference
    round-trips the rotation through cv2.Rodrigues; the residuals themselves are ~10 px)."""
    g = load_golden("ba_reference.npz")
    for name in g["names"]:
        c = _case(g, str(name))
        res = oracle.ba_residuals(**c)
        want = g[str(name) + "_residual"]
        assert res.shape == want.shape
        assert np.abs(res - want).max() < 1e-9, (name, np.abs(res - want).max())
    # global calibration: K and the distortion coefficients ride at the end of the parameter vector (:182-194)
    c = _case(g, "small")
    p = g["small_global_params"]
    calib = p[c["n_cam"] * 7 + c["n_pts"] * 3:]
    c["params"] = p
    c["K4"] = (calib[0], calib[0], calib[1], calib[2])
    c["dist"] = calib[3:]
    assert np.abs(oracle.ba_residuals(**c) - g["small_global_residual"]).max() < 1e-9


def test_kernel_observation_function_on_host_equals_oracle():
    """The kernel's per-observation function (same source, run on the host through iam_debug_ba_host): residual
    against the reference goldens (1e-9 px), analytic Jacobian against central differences of the oracle (1e-5
    relative: the difference quotient's own error)."""
    from imageanalysis_b200 import _capi
    g = load_golden("ba_reference.npz")
    c = _case(g, "small")
    n_cam = c["n_cam"]
    cams = c["params"][:n_cam * 7].reshape(-1, 7)
    pts = c["params"][n_cam * 7:].reshape(-1, 3)
    Jfd = oracle.ba_jacobian_fd(**c)
    want = g["small_residual"].reshape(-1, 2)
    for i in range(len(c["cam_idx"])):
        r, J = _capi.ba_observation_host(cams[c["cam_idx"][i]], pts[c["pt_idx"][i]], c["obs_uv"][i], c["K4"], c["dist"])
        assert np.abs(r - want[i]).max() < 1e-9
        assert (np.abs(J - Jfd[i]) / (1.0 + np.abs(Jfd[i]))).max() < 1e-5


def test_sparsity_pattern_equals_reference_layout():
    """bundle_adjustment_sparsity (optimizer.py:142-169): row 2i / 2i+1 touch the 7 parameters of camera_indices[i]
    and the 3 of point_indices[i]."""
    from imageanalysis_b200 import optimizer
    opt = optimizer.Optimizer()
    cam = np.array([0, 0, 2, 1])
    pt = np.array([3, 0, 1, 1])
    A = opt.bundle_adjustment_sparsity(3, 4, cam, pt).toarray()
    assert A.shape == (8, 3 * 7 + 4 * 3)
    for i in range(4):
        want = np.zeros(33, int)
        want[cam[i] * 7:cam[i] * 7 + 7] = 1
        want[21 + pt[i] * 3:21 + pt[i] * 3 + 3] = 1
        assert (A[2 * i] == want).all() and (A[2 * i + 1] == want).all()


@pytest.mark.gpu
def test_gpu_residual_and_jacobian():
    from imageanalysis_b200 import _capi
    g = load_golden("ba_reference.npz")
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    for name in g["names"]:
        c = _case(g, str(name))
        eng.ba_setup(c["n_cam"], c["n_pts"], c["cam_idx"], c["pt_idx"], c["obs_uv"])
        res, J = eng.ba_eval(c["params"], c["K4"], c["dist"], jac=True)
        assert np.abs(res - g[str(name) + "_residual"]).max() < 1e-9, name
        only = eng.ba_eval(c["params"], c["K4"], c["dist"])                         # residual-only kernel variant
        assert np.abs(only - g[str(name) + "_residual"]).max() < 1e-9, name
        Jfd = oracle.ba_jacobian_fd(**c)
        assert (np.abs(J - Jfd) / (1.0 + np.abs(Jfd))).max() < 1e-5, name
    eng.close()


@pytest.mark.gpu
def test_gpu_optimizer_dropin_in_least_squares():
    """The drop-in Optimizer inside scipy.optimize.least_squares exactly as optimizer.py:491-501 calls it: with the
    analytic Jacobian and with the reference's finite-difference sparsity route the solver reaches the same
    minimum (the cost agrees to 1e-6 relative), and fun() equals the reference residual at x0."""
    from scipy.optimize import least_squares
    from imageanalysis_b200 import optimizer, synth
    prob = synth.ba_problem(n_cam=8, n_pts=300, seed=11, obs_per_cam=120)
    opt = optimizer.Optimizer()
    opt.K, opt.distCoeffs = prob["K"], prob["dist"]
    args = (prob["n_cam"], prob["n_pts"], prob["idx_lists"], prob["uv_lists"])
    f0 = opt.fun(prob["params"], *args)
    cam_idx = np.concatenate([np.full(len(ix), c) for c, ix in enumerate(prob["idx_lists"])])
    pt_idx = np.concatenate(prob["idx_lists"])
    uv = np.concatenate([u.reshape(-1, 2) for u in prob["uv_lists"]])
    K = prob["K"]
    want = oracle.ba_residuals(prob["params"], prob["n_cam"], prob["n_pts"], cam_idx, pt_idx, uv,
                               (K[0, 0], K[1, 1], K[0, 2], K[1, 2]), prob["dist"])
    assert np.abs(f0 - want).max() < 1e-9
    A = opt.bundle_adjustment_sparsity(prob["n_cam"], prob["n_pts"], cam_idx, pt_idx)
    kw = dict(method='trf', loss='linear', ftol=1e-8, x_scale='jac', args=args, max_nfev=40)
    r_an = least_squares(opt.fun, prob["params"], jac=opt.jac, **kw)
    r_fd = least_squares(opt.fun, prob["params"], jac_sparsity=A, **kw)
    assert r_an.cost < 0.2 * 0.5 * float(f0 @ f0)
    assert abs(r_an.cost - r_fd.cost) <= 1e-6 * r_fd.cost


@pytest.mark.gpu
def test_gpu_global_calibration_jacobian():
    """optimize_calib == 'global' (optimizer.py:146-147, :160-166, :181-189): fun() with the calibration riding at the end
    of the vector equals the reference golden, jac() has the declared sparsity pattern, and its eight dense columns
    d/d(f, cu, cv, k1, k2, p1, p2, k3) equal central differences of the CPU oracle (1e-5 relative)."""
    from imageanalysis_b200 import optimizer
    g = load_golden("ba_reference.npz")
    c = _case(g, "small")
    p = np.array(g["small_global_params"], np.float64)
    n_cam, n_pts = c["n_cam"], c["n_pts"]
    n_obs = len(c["cam_idx"])
    idx_lists = [c["pt_idx"][c["cam_idx"] == k] for k in range(n_cam)]
    uv_lists = [c["obs_uv"][c["cam_idx"] == k] for k in range(n_cam)]
    assert (np.diff(c["cam_idx"]) >= 0).all()          # observations are laid out camera by camera (:396-404)
    opt = optimizer.Optimizer()
    opt.optimize_calib = 'global'
    f = opt.fun(p, n_cam, n_pts, idx_lists, uv_lists)
    assert np.abs(f - g["small_global_residual"]).max() < 1e-9
    J = opt.jac(p, n_cam, n_pts, idx_lists, uv_lists)
    A = opt.bundle_adjustment_sparsity(n_cam, n_pts, c["cam_idx"], c["pt_idx"])
    assert J.shape == A.shape == (2 * n_obs, n_cam * 7 + n_pts * 3 + 8)
    assert (J.indices == A.indices).all() and (J.indptr == A.indptr).all()
    Jd = J.toarray()
    base = n_cam * 7 + n_pts * 3

    def residual(q):
        cal = q[base:]
        return oracle.ba_residuals(q, n_cam, n_pts, c["cam_idx"], c["pt_idx"], c["obs_uv"], (cal[0], cal[0], cal[1], cal[2]), cal[3:])

    for k in range(8):
        h = 1e-6 * max(1.0, abs(p[base + k]))
        hi, lo = p.copy(), p.copy()
        hi[base + k] += h
        lo[base + k] -= h
        fd = (residual(hi) - residual(lo)) / (2 * h)
        assert (np.abs(Jd[:, base + k] - fd) / (1.0 + np.abs(fd))).max() < 1e-5, k
    # the camera / point columns are those of the fixed-calibration Jacobian at the same K
    opt2 = optimizer.Optimizer()
    cal = p[base:]
    opt2.K = np.array([[cal[0], 0, cal[1]], [0, cal[0], cal[2]], [0, 0, 1.0]])
    opt2.distCoeffs = cal[3:]
    J2 = opt2.jac(p[:base], n_cam, n_pts, idx_lists, uv_lists).toarray()
    assert np.array_equal(Jd[:, :base], J2)
