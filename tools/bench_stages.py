#!/usr/bin/env python3
"""Stage benchmarks for the rows next to the kNN kernel: the GMS grid filter inside the
match pipeline and the bundle-adjustment residual / Jacobian kernel, each timed with CUDA events on the stream the
library launches on, after warm-up.  One JSON line per stage (also appended to gpurun_out/stages.jsonl).

GMS: whole-pipeline step time with and without the stage (the difference is its cost) on the bench workload.
BA : achieved HBM bandwidth = algorithmic bytes per observation (~104 B read + 176 B written) x
     observations / kernel time, against MEASURED_PEAKS.json hbm_gbs; the reference's CPU path
     (oracle port of Optimizer.fun, numpy) is timed beside it."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=200)
    ap.add_argument("--desc", type=int, default=5000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--ba-cams", type=int, default=1000)
    ap.add_argument("--ba-obs-per-cam", type=int, default=1000)
    args = ap.parse_args()
    import torch
    from imageanalysis_b200 import _capi, synth
    import bench as B
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    out = []

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    # ------------------------------------------------------------------ GMS inside the match pipeline
    des = B.make_frames_gpu(args.frames, args.desc, "SIFT", seed=1234, device=dev)
    rng = np.random.default_rng(7)
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    eng.set_stream(stream.cuda_stream)
    pairs = B.pair_list(args.frames, "sequential")
    # key points: planted rows of neighbouring frames move by a common shift (what the frames' descriptors were
    # built to match), everything else is uniform
    w, h = 5472, 3648
    for f in range(args.frames):
        eng.upload(f, des[f])
        eng.upload_keypoints(f, np.stack([rng.uniform(0, w, args.desc), rng.uniform(0, h, args.desc)], 1).astype(np.float32))
    eng.synchronize()
    res = {}
    for tag, kw in (("no_gms", dict()), ("gms", dict(gms=True, size=(w, h)))):
        prm = _capi.Engine.make_params(**kw)
        res[tag] = timed(lambda: eng.match_pairs_device(pairs, prm), args.steps, args.warmup)
    line = {"stage": "gms_kernel (csrc/gms.cu) inside iam_match_pairs_device", "pairs": len(pairs), "directed_jobs": 2 * len(pairs),
            "ms_per_step_without": res["no_gms"], "ms_per_step_with": res["gms"], "gms_ms_per_step": res["gms"] - res["no_gms"],
            "gms_us_per_directed_job": 1e3 * (res["gms"] - res["no_gms"]) / (2 * len(pairs)),
            "note": "random key points: the filter rejects nearly everything, which is its most expensive path (all four grids x 8 rotations are always evaluated)"}
    out.append(line)
    print(json.dumps(line))
    eng.close()

    # ------------------------------------------------------------------ BA residual + Jacobian
    prob = synth.ba_problem(n_cam=args.ba_cams, n_pts=args.ba_cams * 200, seed=5, obs_per_cam=args.ba_obs_per_cam)
    cam_idx = np.concatenate([np.full(len(ix), c, np.int32) for c, ix in enumerate(prob["idx_lists"])])
    pt_idx = np.concatenate(prob["idx_lists"]).astype(np.int32)
    uv = np.concatenate([u.reshape(-1, 2) for u in prob["uv_lists"]])
    K = prob["K"]
    K4, dist = (K[0, 0], K[1, 1], K[0, 2], K[1, 2]), prob["dist"]
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    eng.set_stream(stream.cuda_stream)
    eng.ba_setup(prob["n_cam"], prob["n_pts"], cam_idx, pt_idx, uv)
    eng.ba_upload_params(prob["params"])
    n_obs = len(cam_idx)
    ms_j = timed(lambda: eng.ba_eval_device(K4, dist, jac=True), 50, 5)
    ms_r = timed(lambda: eng.ba_eval_device(K4, dist, jac=False), 50, 5)
    t0 = time.perf_counter()
    r_h, J_h = eng.ba_eval(prob["params"], K4, dist, jac=True)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak = peaks.get("hbm_gbs") or 6500.0
    bytes_j, bytes_r = 104 + 176, 104 + 16
    from oracle import oracle
    t0 = time.perf_counter()
    sample = min(n_obs, 200000)
    last_cam = int(cam_idx[sample - 1])
    sel = cam_idx <= last_cam
    oracle.ba_residuals(prob["params"], prob["n_cam"], prob["n_pts"], cam_idx[sel], pt_idx[sel], uv[sel], K4, dist)
    cpu_s = time.perf_counter() - t0
    line = {"stage": "ba_kernel (csrc/ba.cu): residual + analytic 2x10 Jacobian blocks", "observations": n_obs,
            "cameras": prob["n_cam"], "points": prob["n_pts"],
            "kernel_ms_residual_and_jacobian": ms_j, "kernel_ms_residual_only": ms_r,
            "observations_per_s": n_obs / (ms_j / 1e3),
            "roofline": {"bound": "hbm", "achieved": n_obs * bytes_j / (ms_j / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": n_obs * bytes_j / (ms_j / 1e3) / 1e9 / peak, "traffic": None,
                         "algorithmic_bytes_per_observation": bytes_j,
                         "residual_only": {"achieved": n_obs * bytes_r / (ms_r / 1e3) / 1e9, "frac": n_obs * bytes_r / (ms_r / 1e3) / 1e9 / peak,
                                           "algorithmic_bytes_per_observation": bytes_r}},
            "e2e_ms_host_params_to_host_jacobian": e2e_ms,
            "cpu_baseline": {"value": int(sel.sum()) / cpu_s, "unit": "observations/s (residual only)", "cores": 1, "kind": "port",
                             "sample": "%d observations of the same problem through oracle.ba_residuals (numpy restatement of Optimizer.fun)" % int(sel.sum())}}
    out.append(line)
    print(json.dumps(line))
    eng.close()

    # ------------------------------------------------------------------ essential-matrix RANSAC (matcher.py:126)
    # BASELINE configs[4]: after matching, filter_by_transform(..., 'essential') per pair.  One two-view scene per
    # pair: 800 matches, 30 % outliers, 0.5 px noise, DJI FC6310S intrinsics; host points in, host masks out.
    Kc = np.array([[3666.5, 0, 2736.0], [0, 3666.5, 1824.0], [0, 0, 1.0]])
    n_pairs, n_m = 1990, 800
    rng = np.random.default_rng(11)
    X = np.c_[rng.uniform(-60, 60, (n_pairs, n_m)).ravel(), rng.uniform(-40, 40, n_pairs * n_m), rng.uniform(60, 90, n_pairs * n_m)]
    base = np.repeat(np.c_[rng.uniform(10, 20, n_pairs), rng.uniform(-3, 3, n_pairs), rng.uniform(-1, 1, n_pairs)], n_m, axis=0)
    yaw = np.repeat(rng.uniform(-0.05, 0.05, n_pairs), n_m)
    X2 = np.c_[np.cos(yaw) * X[:, 0] - np.sin(yaw) * X[:, 1], np.sin(yaw) * X[:, 0] + np.cos(yaw) * X[:, 1], X[:, 2]] - base

    def proj(P):
        return np.c_[Kc[0, 0] * P[:, 0] / P[:, 2] + Kc[0, 2], Kc[1, 1] * P[:, 1] / P[:, 2] + Kc[1, 2]]
    p1 = proj(X) + rng.normal(0, 0.5, (n_pairs * n_m, 2))
    p2 = proj(X2) + rng.normal(0, 0.5, (n_pairs * n_m, 2))
    bad = rng.random(n_pairs * n_m) < 0.3
    p2[bad] = np.c_[rng.uniform(0, 5472, bad.sum()), rng.uniform(0, 3648, bad.sum())]
    off = (np.arange(n_pairs + 1) * n_m).astype(np.int32)
    tol = 5472 ** 0.25
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    eng.set_stream(stream.cuda_stream)
    eng.ransac_pairs(_capi.MODEL_ESSENTIAL, p1, p2, off, Kc, tol)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        mask, E, ninl = eng.ransac_pairs(_capi.MODEL_ESSENTIAL, p1, p2, off, Kc, tol)
    gpu_ms = (time.perf_counter() - t0) * 1e3 / reps
    cpu = None
    try:
        import cv2
        cv2.setNumThreads(os.cpu_count())
        t0 = time.perf_counter()
        agree = []
        n_cpu = 16
        for s in range(n_cpu):
            a, b = p1[off[s]:off[s + 1]].astype(np.float32), p2[off[s]:off[s + 1]].astype(np.float32)
            _, m = cv2.findEssentialMat(a, b, Kc, cv2.RANSAC, threshold=tol)
            m = m.ravel().astype(bool)
            g = mask[off[s]:off[s + 1]].astype(bool)
            agree.append((m & g).sum() / max(1, (m | g).sum()))
        cpu_ms = (time.perf_counter() - t0) * 1e3 / n_cpu
        cpu = {"value": 1e3 / cpu_ms, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "reference",
               "sample": "cv2.findEssentialMat(p1, p2, K, cv2.RANSAC, threshold=tol) on %d of the pairs" % n_cpu,
               "inlier_iou_vs_gpu_min": float(min(agree)), "inlier_iou_vs_gpu_mean": float(np.mean(agree))}
    except ImportError:
        pass
    line = {"stage": "ransac_kernel (csrc/ransac.cu): 5-point essential-matrix RANSAC, host points in -> host masks out",
            "pairs": n_pairs, "matches_per_pair": n_m, "outlier_fraction": 0.3, "ms_per_call": gpu_ms,
            "pairs_per_s": n_pairs / (gpu_ms / 1e3), "mean_inliers": float(ninl.mean()), "cpu_baseline": cpu}
    out.append(line)
    print(json.dumps(line))
    # ------------------------------------------------------------------ ORB detect + describe (image.py:243-245, :324)
    try:
        import cv2
        from imageanalysis_b200 import detector
        rng = np.random.default_rng(3)
        img = cv2.GaussianBlur(rng.integers(0, 256, (1459, 2189)).astype(np.uint8), (0, 0), 1.5)   # 0.4 x (5472 x 3648)
        for nfeat in (5000, 20000):
            for _ in range(3):
                r = detector.orb_detect_and_compute(img, nfeat)
            t0 = time.perf_counter()
            reps = 10
            for _ in range(reps):
                r = detector.orb_detect_and_compute(img, nfeat)
            gpu_ms = (time.perf_counter() - t0) * 1e3 / reps
            orb = cv2.ORB_create(nfeat)
            orb.detectAndCompute(img, None)
            t0 = time.perf_counter()
            for _ in range(3):
                kps, des = orb.detectAndCompute(img, None)
            cpu_ms = (time.perf_counter() - t0) * 1e3 / 3
            want = {(round(k.pt[0], 2), round(k.pt[1], 2), k.octave): bytes(d) for k, d in zip(kps, des)}
            got = {(round(float(p[0]), 2), round(float(p[1]), 2), int(o)): bytes(d) for p, o, d in zip(r["pt"], r["octave"], r["des"])}
            same = sum(1 for k in want if got.get(k) == want[k])
            line = {"stage": "ORB detect + describe (csrc/orb.cu), grey host image in -> host key points + descriptors out",
                    "image": "2189 x 1459 (the reference's 0.4 scale of a 5472 x 3648 frame)", "nfeatures": nfeat,
                    "ms_per_frame": gpu_ms, "frames_per_s": 1e3 / gpu_ms, "keypoints": int(len(r["pt"])),
                    "identical_to_cv2": {"keypoints_cv2": len(want), "same_position_octave_and_descriptor": same},
                    "cpu_baseline": {"value": 1e3 / cpu_ms, "unit": "frames/s", "cores": cv2.getNumThreads(), "kind": "reference",
                                     "sample": "cv2.ORB_create(%d).detectAndCompute on the same frame, mean of 3" % nfeat}}
            out.append(line)
            print(json.dumps(line))
    except ImportError:
        pass
    # ------------------------------------------------------------------ SIFT detect + describe (image.py:236-237, :324)
    sys.path.insert(0, ROOT)
    import bench as B
    line = dict(stage="SIFT detect + describe (csrc/sift.cu), grey host image in -> host key points + descriptors out", **B.sift_leg(0))
    out.append(line)
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "stages.jsonl"), "w") as f:
        for l in out:
            f.write(json.dumps(l) + "\n")


if __name__ == "__main__":
    main()
