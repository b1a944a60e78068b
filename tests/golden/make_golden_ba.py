#!/usr/bin/env python3
"""Golden vectors for the bundle-adjustment residual (scripts/lib/optimizer.py:174-279, Optimizer.fun).

The reference's Optimizer is imported UNMODIFIED from /root/reference (shims for props / matplotlib as in
make_golden.py) and its fun() -- quaternion_matrix + cv2.Rodrigues + cv2.projectPoints per camera -- is evaluated on
synthetic survey problems.  Recorded: the parameter vector, the per-camera observation lists, K, the distortion
coefficients and the residual vector fun() returned.

usage: python tests/golden/make_golden_ba.py      (from the repo root; needs /root/reference and cv2)
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, "/root/reference/scripts/lib/archive")
sys.path.insert(0, "/root/reference/scripts")
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from lib import optimizer  # noqa: E402  the unmodified reference module

from imageanalysis_b200 import synth  # noqa: E402


def run_reference(params, n_cam, n_pts, idx_lists, uv_lists, K, dist, calib="none"):
    opt = optimizer.Optimizer("/tmp")
    opt.K = K
    opt.distCoeffs = dist
    opt.optimize_calib = calib
    opt.camera_map_fwd = {i: i for i in range(n_cam)}
    with contextlib.redirect_stdout(io.StringIO()):
        return np.asarray(opt.fun(params, n_cam, n_pts, idx_lists, uv_lists), np.float64)


def main():
    out = {}
    cases = {"small": dict(n_cam=6, n_pts=120, seed=1, obs_per_cam=60),
             "strip": dict(n_cam=40, n_pts=3000, seed=2, obs_per_cam=200),
             "empty_cams": dict(n_cam=9, n_pts=200, seed=3, obs_per_cam=50, empty=(0, 4, 8))}
    for name, kw in cases.items():
        prob = synth.ba_problem(**kw)
        res = run_reference(prob["params"], prob["n_cam"], prob["n_pts"], prob["idx_lists"], prob["uv_lists"], prob["K"], prob["dist"])
        print(name, "observations", len(res) // 2, "mre", float(np.mean(np.abs(res))))
        out[name + "_params"] = prob["params"]
        out[name + "_n"] = np.int64([prob["n_cam"], prob["n_pts"]])
        out[name + "_K"] = prob["K"]
        out[name + "_dist"] = prob["dist"]
        out[name + "_cam_idx"] = np.concatenate([np.full(len(ix), c, np.int32) for c, ix in enumerate(prob["idx_lists"])])
        out[name + "_pt_idx"] = np.concatenate([np.asarray(ix, np.int32) for ix in prob["idx_lists"]])
        out[name + "_uv"] = np.concatenate([np.asarray(u, np.float64).reshape(-1, 2) for u in prob["uv_lists"]])
        out[name + "_residual"] = res
        if name == "small":   # global calibration mode: K and the distortion ride at the end of the vector (:182-194)
            calib = np.array([prob["K"][0, 0] * 1.01, prob["K"][0, 2] - 3.0, prob["K"][1, 2] + 2.0, -0.02, 0.01, 0.0005, -0.0003, 0.002])
            p2 = np.concatenate([prob["params"], calib])
            out["small_global_params"] = p2
            out["small_global_residual"] = run_reference(p2, prob["n_cam"], prob["n_pts"], prob["idx_lists"], prob["uv_lists"],
                                                         prob["K"], prob["dist"], calib="global")
    out["names"] = np.array(sorted(cases))
    np.savez_compressed(os.path.join(HERE, "ba_reference.npz"), **out)


if __name__ == "__main__":
    main()
