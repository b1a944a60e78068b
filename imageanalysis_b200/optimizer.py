"""Drop-in for the residual path of the reference's `lib.optimizer.Optimizer`
(scripts/lib/optimizer.py): `fun()` (:174-279) evaluated for all cameras in ONE
CUDA launch instead of a Python loop of cv2.projectPoints calls, and `jac()`,
the analytic Jacobian of it in the sparsity pattern of
`bundle_adjustment_sparsity()` (:142-169), which the reference leaves to
SciPy's finite differences (least_squares(..., jac_sparsity=A), :491-501).

Same call signatures, so `least_squares(opt.fun, x0, jac=opt.jac, ...)` (or the
reference's own call with jac_sparsity) works unchanged:

    opt = Optimizer(root)
    opt.K, opt.distCoeffs = K, dist                      # as Optimizer.setup() leaves them (:294-300)
    err = opt.fun(params, n_cameras, n_points, by_camera_point_indices, by_camera_points_2d)
    J   = opt.jac(params, n_cameras, n_points, by_camera_point_indices, by_camera_points_2d)   # scipy.sparse.csr_matrix

cam_method is 'ned_quat' (the reference's setting, :84-85).  Project set-up, bounds, re-centring and file output of the
reference class are outside the accelerated path.  No CPU fallback: without libiamatch.so / a B200 the calls raise.
"""
from __future__ import annotations

import numpy as np

from . import _capi

try:
    from .logger import log  # type: ignore
except ImportError:
    def log(*args):
        print(*args)


class Optimizer():
    def __init__(self, root=None):
        self.root = root
        self.last_mre = None
        self.optimize_calib = 'none'      # 'global': K and distCoeffs ride at the end of the vector (:182-194)
        self.cam_method = 'ned_quat'
        self.ncp = 7
        self.K = None
        self.distCoeffs = None
        self.cam2body = np.array([[0, 0, 1], [1, 0, 0], [0, 1, 0]], dtype=float)
        self.body2cam = np.linalg.inv(self.cam2body)
        self._eng = None
        self._key = None
        self._shape = None

    # -- problem structure -> device (once per problem) -----------------------
    def _engine(self):
        if self._eng is None:
            from . import matcher as _matcher          # one process per GPU: the device configure() uses
            self._eng = _capi.Engine(_capi.NORM_L2, 128, int(getattr(_matcher, "device", 0)))
        return self._eng

    def rebind(self):
        """Forget the resident problem structure (the next fun()/jac() uploads it again)."""
        self._key = None

    def _bind(self, n_cameras, n_points, by_camera_point_indices, by_camera_points_2d):
        cam_idx, pt_idx, uv = [], [], []
        for i in range(n_cameras):                 # cameras without observations are skipped (:203-204)
            idx = np.asarray(by_camera_point_indices[i], np.int64).ravel()
            if len(idx) == 0:
                continue
            cam_idx.append(np.full(len(idx), i, np.int32))
            pt_idx.append(idx.astype(np.int32))
            uv.append(np.asarray(by_camera_points_2d[i], np.float64).reshape(len(idx), 2))
        cam_idx = np.concatenate(cam_idx) if cam_idx else np.zeros(0, np.int32)
        pt_idx = np.concatenate(pt_idx) if pt_idx else np.zeros(0, np.int32)
        uv = np.concatenate(uv) if uv else np.zeros((0, 2))
        # The resident structure is keyed on its CONTENT (ids of the caller's lists can be reused by CPython once a
        # previous problem is freed, and the lists can be edited in place): a hash of the flattened arrays.
        import hashlib
        h = hashlib.blake2b(digest_size=16)
        for a in (cam_idx, pt_idx, np.ascontiguousarray(uv)):
            h.update(a.tobytes())
        key = (h.digest(), n_cameras, n_points)
        if key == self._key:
            return
        self._engine().ba_setup(n_cameras, n_points, cam_idx, pt_idx, uv)
        self.camera_indices, self.point_indices = cam_idx, pt_idx
        self._key = key
        self._shape = (n_cameras, n_points, len(cam_idx))

    def _calib(self, params, n_cameras, n_points):
        if self.optimize_calib == 'global':
            c = np.asarray(params, np.float64)[n_cameras * self.ncp + n_points * 3:]
            return (c[0], c[0], c[1], c[2]), c[3:8]
        K = np.asarray(self.K, np.float64)
        d = np.zeros(5)                                   # (k1, k2, p1, p2, k3); shorter lists are zero-padded as cv2 does
        dc = np.asarray(self.distCoeffs if self.distCoeffs is not None else [], np.float64).ravel()[:5]
        d[:len(dc)] = dc
        return (K[0, 0], K[1, 1], K[0, 2], K[1, 2]), d

    # -- optimizer.py:174-279 --------------------------------------------------
    def fun(self, params, n_cameras, n_points, by_camera_point_indices, by_camera_points_2d):
        self._bind(n_cameras, n_points, by_camera_point_indices, by_camera_points_2d)
        K4, dist = self._calib(params, n_cameras, n_points)
        error = self._engine().ba_eval(params, K4, dist)
        mre = float(np.mean(np.abs(error))) if len(error) else 0.0
        if self.last_mre is None or 1.0 - mre / self.last_mre > 0.001:    # runtime feedback as :246-250
            self.last_mre = mre
            log('mre: %.3f std: %.3f max: %.2f' % (mre, float(np.std(error)), float(np.amax(np.abs(error)))))
        return error

    def jac(self, params, n_cameras, n_points, by_camera_point_indices, by_camera_points_2d):
        """d fun / d params as scipy.sparse.csr_matrix of shape (2*n_obs, n_cameras*7 + n_points*3 [+ 8]): the pattern
        bundle_adjustment_sparsity() declares, filled analytically -- in global-calibration mode including the eight
        dense columns d/d(f, cu, cv, k1, k2, p1, p2, k3) at the end (:160-166)."""
        from scipy.sparse import csr_matrix
        self._bind(n_cameras, n_points, by_camera_point_indices, by_camera_points_2d)
        K4, dist = self._calib(params, n_cameras, n_points)
        _, J = self._engine().ba_eval(params, K4, dist, jac=True)
        n_obs = self._shape[2]
        glob = self.optimize_calib == 'global'
        per = 18 if glob else 10
        n_cols = n_cameras * 7 + n_points * 3 + (8 if glob else 0)
        cols = np.empty((n_obs, per), np.int64)
        cols[:, :7] = self.camera_indices[:, None] * 7 + np.arange(7)
        cols[:, 7:10] = n_cameras * 7 + self.point_indices[:, None] * 3 + np.arange(3)
        if glob:
            cols[:, 10:] = n_cameras * 7 + n_points * 3 + np.arange(8)
            J = np.concatenate([J, self._engine().ba_calib_jacobian(K4, dist)], axis=2)
        indices = np.repeat(cols, 2, axis=0).ravel()            # rows 2i and 2i+1 share their columns
        indptr = np.arange(0, 2 * per * n_obs + 1, per)
        return csr_matrix((J.ravel(), indices, indptr), shape=(2 * n_obs, n_cols))

    # -- optimizer.py:142-169 (same pattern, built without the Python loops) ---
    def bundle_adjustment_sparsity(self, n_cameras, n_points, camera_indices, point_indices):
        from scipy.sparse import csr_matrix
        camera_indices = np.asarray(camera_indices, np.int64)
        point_indices = np.asarray(point_indices, np.int64)
        m = camera_indices.size * 2
        n = n_cameras * self.ncp + n_points * 3 + (8 if self.optimize_calib == 'global' else 0)
        per = self.ncp + 3 + (8 if self.optimize_calib == 'global' else 0)
        cols = np.empty((camera_indices.size, per), np.int64)
        cols[:, :self.ncp] = camera_indices[:, None] * self.ncp + np.arange(self.ncp)
        cols[:, self.ncp:self.ncp + 3] = n_cameras * self.ncp + point_indices[:, None] * 3 + np.arange(3)
        if self.optimize_calib == 'global':
            cols[:, self.ncp + 3:] = n_cameras * self.ncp + n_points * 3 + np.arange(8)
        indices = np.repeat(cols, 2, axis=0).ravel()
        indptr = np.arange(0, per * m + 1, per)
        return csr_matrix((np.ones(indices.size, dtype=int), indices, indptr), shape=(m, n))
