#!/usr/bin/env python3
"""bench.py — image-pairs matched per second on BASELINE.json's configs.

A "step" is one pass of the hot path over one batch: every candidate pair goes through
kNN (both directions) -> metric reduction -> cross-check -> per-pair match table.

  --gpus 1 (default) : BASELINE configs[1] -- a 500-frame strip, 5000 SIFT descriptors per frame, the reference's
                       live pair generator |i-j| <= 4 (matcher.py:899) -> 1990 pairs           (--workload strip)
  --gpus N > 1       : BASELINE configs[3] -- 2812 frames on the 38 x 74 serpentine survey grid, the reference's
                       camera-distance pair filter (matcher.py:858-903) -> 42 694 pairs, the pair list sharded in
                       contiguous blocks across the ranks (strong scaling), every rank uploads only the frames its
                       block touches, ONE compact all-gather collects the match tables        (--workload bates)

  value : pairs/s with descriptors already resident in HBM (device timed, CUDA events, max over ranks)
  e2e   : the same job through the public one-call API with HOST buffers: H2D of every frame's float32
          descriptors from pinned memory + layout conversion + matching + D2H of the match tables (+ the
          gather for N > 1), every step
  roofline : algorithmic 2*N*M*128 FLOP per pair / the kNN kernel's own time, against the measured peak of the
          MMA kind the kernel issues
  parity_spot : after the timed region, three pairs of the run are compared with the CPU oracle
  orb / sift_detect : short extra legs of the default N = 1 line -- BASELINE configs[2] (ORB, Hamming) and the SIFT
          detect + describe stage in front of the matcher (one survey-sized frame, cv2 timed and compared beside it)
  cpu_baseline / --impl reference : the reference's CPU path (cv2.BFMatcher both directions + the Python
          reduction of matcher.py:253-269 + cross-check) on the box's host cores, on a bounded sample of the
          SAME frames and pair list
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PAIR_L2 = 2.0 * 5000 * 5000 * 128      # one N x M product per pair
OP_PER_PAIR_HAMMING = 2.0 * 5000 * 5000 * 256
SEED = 1234


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=0, help="0: 500 (strip) / 2812 (bates)")
    ap.add_argument("--desc", type=int, default=5000)
    ap.add_argument("--detector", default="SIFT", choices=["SIFT", "ORB"])
    ap.add_argument("--pairs", default="sequential", choices=["sequential", "all"])
    ap.add_argument("--workload", default="auto", choices=["auto", "strip", "bates", "pipeline"],
                    help="auto: strip (BASELINE configs[1]) on one GPU, bates (configs[3], pair-sharded) on several; "
                         "strip with --gpus N > 1 = N independent replicas (weak scaling); pipeline: BASELINE configs[4] -- "
                         "the bates project end to end: match + essential-matrix RANSAC per pair + bundle-adjustment "
                         "residual/Jacobian evaluation")
    ap.add_argument("--engine", default="auto", choices=["auto", "umma", "umma_f16", "simt"])
    ap.add_argument("--gather", default="packed", choices=["packed", "padded"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-orb", action="store_true", help="skip the short ORB (configs[2]) leg of the default line")
    ap.add_argument("--no-spot", action="store_true")
    ap.add_argument("--cpu-pairs", type=int, default=4)
    return ap.parse_args()


# ----------------------------------------------------------------------------- data
def make_frames_gpu(frames, n, detector, seed, device, keep=None):
    """Synthetic descriptors with OpenCV-SIFT statistics (see imageanalysis_b200/synth.py), generated with torch
    (on `device`; the stream of frames is a chain -- 40 % of every frame's rows are noisy copies of rows of the
    previous frame -- so frame f is the same whatever the total count).  Returns {frame: uint8 [n, D]} on the HOST
    for the frames in `keep` (all when None)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = {}
    prev = None
    last = frames - 1 if keep is None else max(keep)
    for f in range(last + 1):
        if detector == "SIFT":
            v = torch._standard_gamma(torch.full((n, 128), 0.6, device=device), generator=g)
            v = v / v.norm(dim=1, keepdim=True).clamp_min(1e-12)
            v = v.clamp_max(0.2)
            v = v / v.norm(dim=1, keepdim=True).clamp_min(1e-12)
            d = (v * 512.0).round().clamp(0, 255).to(torch.uint8)
            if prev is not None:
                m = int(0.4 * n)
                src = torch.randperm(n, device=device, generator=g)[:m]
                dst = torch.randperm(n, device=device, generator=g)[:m]
                noise = torch.randint(-3, 4, (m, 128), device=device, generator=g)
                d[dst] = (prev[src].to(torch.int32) + noise).clamp(0, 255).to(torch.uint8)
        else:
            d = torch.randint(0, 256, (n, 32), device=device, generator=g, dtype=torch.uint8)
            if prev is not None:
                m = int(0.4 * n)
                src = torch.randperm(n, device=device, generator=g)[:m]
                dst = torch.randperm(n, device=device, generator=g)[:m]
                flips = torch.zeros((m, 32), dtype=torch.uint8, device=device)
                for _ in range(12):
                    byte = torch.randint(0, 32, (m,), device=device, generator=g)
                    bit = torch.randint(0, 8, (m,), device=device, generator=g)
                    flips[torch.arange(m, device=device), byte] ^= (1 << bit).to(torch.uint8)
                d[dst] = prev[src] ^ flips
        prev = d
        if keep is None or f in keep:
            out[f] = d.cpu().numpy()
    return out


def pair_list(frames, mode):
    if mode == "all":
        return np.asarray([(i, j) for i in range(frames) for j in range(i + 1, frames)], np.int32)
    return np.asarray([(i, j) for i in range(frames) for j in range(i + 1, min(frames, i + 5))], np.int32)


def bates_pairs(frames):
    """BASELINE configs[3]: 38 flight lines x 74 frames, 15 m along-track, 25 m cross-track,
    serpentine; pairs = the reference's camera-distance window (matcher.py:858-903, pairs.worklist 'geotag')."""
    from imageanalysis_b200 import pairs as wl, synth
    neds = synth.survey_grid_neds()[:frames]
    return wl.pair_array(wl.worklist(neds, "geotag"))


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU reference path
def cpu_pairs_per_s(des_u8, pairs, detector, n_pairs, repeats=1, context=False):
    """The reference's CPU path for the first `n_pairs` pairs of `pairs` on the frames `des_u8` ({frame: u8 rows},
    the SAME frames the GPU arm matches): knnMatch both ways with the exact matcher (cv2.BFMatcher when cv2 is
    importable: kind 'reference'; otherwise the C restatement oracle/oracle_knn.c: kind 'port') plus the Python
    reduction of matcher.py:253-269 and the cross-check (:187-200)."""
    from oracle import oracle
    norm = oracle.NORM_L2 if detector == "SIFT" else oracle.NORM_HAMMING
    max_distance = 270.0 if detector == "SIFT" else 64.0
    cores = os.cpu_count() or 1
    try:
        import cv2
        cv2.setNumThreads(cores)
        bf = cv2.BFMatcher(cv2.NORM_L2 if detector == "SIFT" else cv2.NORM_HAMMING)
        kind, cores_used = "reference", cv2.getNumThreads()

        def knn(a, b):
            m = bf.knnMatch(a, b, k=2)
            idx = np.int32([[x[0].trainIdx, x[1].trainIdx] for x in m])
            dist = np.float32([[x[0].distance, x[1].distance] for x in m])
            return idx, dist
        conv = (lambda d: d.astype(np.float32)) if detector == "SIFT" else (lambda d: d)
    except ImportError:
        kind, cores_used = "port", cores
        knn = lambda a, b: oracle.knn(a, b, 2, norm, threads=cores)   # noqa: E731
        conv = lambda d: d                                           # noqa: E731
    sample = pairs[:n_pairs]
    used = sorted({int(i) for p in sample for i in p})
    host = {i: conv(des_u8[i]) for i in used}
    best = None
    for _ in range(repeats + 1):            # first round is the warm-up
        t0 = time.perf_counter()
        for i, j in sample:
            i1, d1 = knn(host[int(i)], host[int(j)])
            p1 = oracle.reduce_ref_metric(i1, d1, 0.75, max_distance, 2000, 25)
            if len(p1) >= 25:
                i2, d2 = knn(host[int(j)], host[int(i)])
                p2 = oracle.reduce_ref_metric(i2, d2, 0.75, max_distance, 2000, 25)
            else:
                p2 = []
            oracle.filter_cross_check(p1, p2)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    out = {"value": len(sample) / best, "unit": "pairs/s", "cores": cores_used, "kind": kind,
           "sample": "the first %d of the %d pairs of this workload (same synthetic frames as the GPU arm; every pair "
                     "costs the same on the CPU path, so the rate extrapolates), both kNN directions + matcher.py:253-269 "
                     "reduction + cross-check, best of %d timed rounds after one warm-up round" % (
                         len(sample), len(pairs), max(1, repeats))}
    if context and kind == "reference":
        # Context only: the same pair on ONE host thread, and the matcher the reference literally
        # configures -- an approximate FLANN kd-tree search (matcher.py:62-79, checks=100), kNN both directions
        # without the reduction.
        try:
            import cv2
            i, j = (int(x) for x in sample[0])
            a, b = host[i], host[j]
            cv2.setNumThreads(1)
            t0 = time.perf_counter()
            knn(a, b)
            knn(b, a)
            one = time.perf_counter() - t0
            cv2.setNumThreads(cores)
            ctx = {"one_thread_knn_pairs_per_s": 1.0 / one}
            if detector == "SIFT":
                fl = cv2.FlannBasedMatcher(dict(algorithm=1, trees=5), dict(checks=100))
                t0 = time.perf_counter()
                fl.knnMatch(a, b, k=2)
                fl.knnMatch(b, a, k=2)
                ctx["flann_kdtree_checks100_knn_pairs_per_s"] = 1.0 / (time.perf_counter() - t0)
            out["context"] = ctx
        except Exception as e:  # noqa: BLE001  (context figures must never break the bench line)
            out["context"] = {"error": str(e)[:80]}
    return out


def resolve_workload(args, world):
    if args.workload == "auto":
        args.workload = "strip" if world == 1 else "bates"
    if not args.frames:
        args.frames = 2812 if args.workload in ("bates", "pipeline") else 500
    return args.workload


def workload_config(args, n_pairs_total, world, pairs_per_gpu=None, frames_per_gpu=None):
    n_pad = -(-args.desc // 256) * 256
    row_bytes = 192 if args.detector == "SIFT" else 288 * 2
    if args.workload == "bates":
        return {"workload": "%d frames (38 x 74 serpentine survey grid) x %d %s descriptors/frame, geotag-neighbour pair "
                            "list (reference matcher.py:858-903), %d pairs in all, sharded across %d GPU(s)" % (
                                args.frames, args.desc, args.detector, n_pairs_total, world),
                "frames": args.frames, "desc_per_frame": args.desc, "pairs_total": n_pairs_total,
                "pairs_per_gpu": pairs_per_gpu or -(-n_pairs_total // world), "frames_per_gpu": frames_per_gpu,
                "match_ratio": 0.75, "cap": 2000, "min_pairs": 25,
                "l2_hygiene": "inputs larger than L2 (operand forms %.2f GB per GPU)" % (
                    (frames_per_gpu or args.frames) * n_pad * row_bytes / 1e9),
                "parallelism": ("pair-sharded x%d (contiguous blocks of the (i, j)-sorted work list; a rank holds only the "
                                "frames its block touches) + 1 NCCL all-gather of the compact match tables" % world)
                if world > 1 else "single GPU"}
    return {"workload": "%d frames x %d %s descriptors/frame, %s pair list (%d pairs) per GPU" % (
                args.frames, args.desc, args.detector,
                "|i-j|<=4 (reference matcher.py:899)" if args.pairs == "sequential" else "all-pairs",
                n_pairs_total // world),
            "frames_per_gpu": args.frames, "desc_per_frame": args.desc, "pairs_per_gpu": n_pairs_total // world,
            "pairs_total": n_pairs_total, "match_ratio": 0.75, "cap": 2000, "min_pairs": 25,
            "l2_hygiene": "inputs larger than L2 (operand forms %.2f GB per GPU)" % (args.frames * n_pad * row_bytes / 1e9),
            "parallelism": "%d independent replicas + 1 NCCL all-gather of the match tables" % world if world > 1 else "single GPU"}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path (cv2.BFMatcher both directions + the
    reduction and cross-check of scripts/lib/matcher.py), all host threads, on a bounded sample (the first
    --cpu-pairs pairs) of the SAME frames and pair list the GPU arm runs.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    resolve_workload(args, world)
    import torch
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))) if torch.cuda.is_available() else torch.device("cpu")
    all_pairs = bates_pairs(args.frames) if args.workload == "bates" else pair_list(args.frames, args.pairs)
    sample = all_pairs[:args.cpu_pairs]
    keep = {int(i) for p in sample for i in p}
    # the generator is a chain: the first frames are identical to the GPU arm's (same seed, same device kind)
    des = make_frames_gpu(max(keep) + 1, args.desc, args.detector, seed=SEED, device=dev, keep=keep)
    steps = []
    base = None
    for s in range(args.warmup + args.steps):
        base = cpu_pairs_per_s(des, all_pairs, args.detector, args.cpu_pairs, repeats=0 if s else 1,
                               context=(s == args.warmup + args.steps - 1))
        if s >= args.warmup:
            steps.append(base["value"])
    v = statistics.median(steps)
    base["value"] = v
    P_total = len(all_pairs) * (world if args.workload == "strip" else 1)
    line = {"impl": "reference", "metric": "image-pairs matched/sec (5000 %s desc/img)" % args.detector, "value": v,
            "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * args.cpu_pairs / v, "higher_is_better": True,
            "scaling": "strong" if args.workload == "bates" else "weak",
            "vs_baseline": None, "dtype": "f32" if args.detector == "SIFT" else "u8", "data": "synthetic",
            "config": workload_config(args, P_total, world),
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- parity spot check
def parity_spot(des_u8, pairs, table, count, detector, which):
    """Untimed: the match tables of a few pairs of THIS run against the CPU oracle (bit-exact index lists)."""
    from oracle import oracle
    norm = oracle.NORM_L2 if detector == "SIFT" else oracle.NORM_HAMMING
    maxd = 270.0 if detector == "SIFT" else 64.0
    bad = []
    for p in which:
        a, b = int(pairs[p][0]), int(pairs[p][1])
        f, _ = oracle.bidirectional(des_u8[a], des_u8[b], norm, 0.75, maxd, threads=os.cpu_count() or 4)
        if table[p, :count[p]].tolist() != f:
            bad.append(int(p))
    return {"status": "ok" if not bad else "MISMATCH", "pairs_checked": [[int(pairs[p][0]), int(pairs[p][1])] for p in which],
            "rows_checked": int(sum(int(count[p]) for p in which)), "mismatching_pairs": bad,
            "checker": "oracle.bidirectional (CPU restatement of matcher.py:218-318), whole tables compared"}



# ----------------------------------------------------------------------------- BASELINE configs[4]: the whole project
PIPE_K = np.array([[3666.666504, 0.0, 2736.0], [0.0, 3666.666504, 1824.0], [0.0, 0.0, 1.0]])   # cameras/DJI_FC6310S.json
PIPE_W, PIPE_H = 5472, 3648


def make_project_gpu(frames, n, seed, device, keep, ba_per_frame=175):
    """The survey as a consistent 3-D scene: every descriptor row of a frame belongs to a ground point, a planted row
    (40 % of a frame, copied with noise from the previous frame, as make_frames_gpu does) inherits the point of the row
    it was copied from, every other row gets a new point inside the frame's footprint; key points = the projection of
    the points through the frame's nadir camera (survey_grid_neds pose, a few degrees of yaw) + 0.5 px noise.  So the
    matches the kNN finds between ANY two frames obey one two-view geometry each, which is what the essential-matrix
    RANSAC stage needs.  Also returns the bundle-adjustment problem of the shape Optimizer.setup() builds
    (optimizer.py:283-420) from `ba_per_frame` inherited points per consecutive frame pair (two observations each).
    Returns (des {f: u8 [n,128]}, uv {f: f32 [n,2]}, ba dict)."""
    import torch
    from imageanalysis_b200 import synth
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f64 = torch.float64
    neds = torch.tensor(synth.survey_grid_neds()[:frames], dtype=f64, device=device)
    K = torch.tensor(PIPE_K, dtype=f64, device=device)
    IK = torch.linalg.inv(K)
    cam2body = torch.tensor([[0.0, 0, 1], [1, 0, 0], [0, 1, 0]], dtype=f64, device=device)
    yaw = (torch.rand(frames, generator=g, device=device, dtype=f64) - 0.5) * np.deg2rad(6.0)
    pitch = torch.full((frames,), -np.pi / 2, dtype=f64, device=device)
    cy, sy, cp, sp = torch.cos(yaw / 2), torch.sin(yaw / 2), torch.cos(pitch / 2), torch.sin(pitch / 2)
    quat = torch.stack([cp * cy, -sp * sy, sp * cy, cp * sy], 1)       # body -> ned, ZYX euler (yaw, pitch, roll = 0)

    def b2n(q):
        w, x, y, z = q
        return torch.stack([torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)]),
                            torch.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)]),
                            torch.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)])])
    des_out, uv_out = {}, {}
    prev = prev_X = prev_uv = None
    ba_cam, ba_uv, ba_X = [], [], []
    last = max(keep)
    for f in range(last + 1):
        v = torch._standard_gamma(torch.full((n, 128), 0.6, device=device), generator=g)
        v = v / v.norm(dim=1, keepdim=True).clamp_min(1e-12)
        v = v.clamp_max(0.2)
        v = v / v.norm(dim=1, keepdim=True).clamp_min(1e-12)
        d = (v * 512.0).round().clamp(0, 255).to(torch.uint8)
        R_b2n = b2n(quat[f])
        # new ground points: a uniformly random pixel's ray, cut at rough ground (D = N(0, 1.5 m))
        pix = torch.stack([torch.rand(n, generator=g, device=device, dtype=f64) * PIPE_W,
                           torch.rand(n, generator=g, device=device, dtype=f64) * PIPE_H, torch.ones(n, dtype=f64, device=device)], 1)
        ray = pix @ (R_b2n @ cam2body @ IK).T
        ground = torch.randn(n, generator=g, device=device, dtype=f64) * 1.5
        t = (ground - neds[f, 2]) / ray[:, 2]
        X = neds[f] + ray * t[:, None]
        src = dst = None
        if prev is not None:
            m = int(0.4 * n)
            src = torch.randperm(n, device=device, generator=g)[:m]
            dst = torch.randperm(n, device=device, generator=g)[:m]
            noise = torch.randint(-3, 4, (m, 128), device=device, generator=g)
            d[dst] = (prev[src].to(torch.int32) + noise).clamp(0, 255).to(torch.uint8)
            X[dst] = prev_X[src]
        Rc = cam2body.T @ R_b2n.T                                       # ned -> camera (Image.get_proj, image.py:543-553)
        Xc = (X - neds[f]) @ Rc.T
        uvh = Xc @ K.T
        uv = uvh[:, :2] / uvh[:, 2:3] + torch.randn(n, 2, generator=g, device=device, dtype=f64) * 0.5
        if src is not None and ba_per_frame > 0:                       # two observations of each of the first inherited points
            k = ba_per_frame
            ba_cam.append(torch.stack([torch.full((k,), f - 1, device=device), torch.full((k,), f, device=device)], 1))
            ba_uv.append(torch.stack([prev_uv[src[:k]], uv[dst[:k]]], 1))
            ba_X.append(X[dst[:k]])
        prev, prev_X, prev_uv = d, X, uv
        if f in keep:
            des_out[f] = d.cpu().numpy()
            uv_out[f] = uv.to(torch.float32).cpu().numpy()
    ba = None
    if ba_cam:
        cam = torch.cat(ba_cam).cpu().numpy()            # [n_pts, 2]
        uvs = torch.cat(ba_uv).cpu().numpy()             # [n_pts, 2, 2]
        Xs = torch.cat(ba_X).cpu().numpy()
        n_pts = len(Xs)
        cam_idx = cam.reshape(-1).astype(np.int32)
        pt_idx = np.repeat(np.arange(n_pts, dtype=np.int32), 2)
        obs = uvs.reshape(-1, 2)
        order = np.argsort(cam_idx, kind="stable")       # by camera, then list order (optimizer.py:396-404)
        cams7 = np.concatenate([neds.cpu().numpy(), quat.cpu().numpy()], 1)[:last + 1]
        rng = np.random.default_rng(seed)
        params = np.concatenate([cams7.ravel(), (Xs + rng.normal(0, 0.3, Xs.shape)).ravel()])
        ba = dict(n_cam=last + 1, n_pts=n_pts, cam_idx=cam_idx[order], pt_idx=pt_idx[order], uv=obs[order], params=params,
                  K4=(PIPE_K[0, 0], PIPE_K[1, 1], PIPE_K[0, 2], PIPE_K[1, 2]), dist=np.zeros(5))
    return des_out, uv_out, ba


def run_pipeline(args):
    """BASELINE configs[4]: 2812 frames end to end -- match (kNN both ways + reduction + cross-check) ->
    filter_by_transform(..., 'essential') for every pair on the device tables (matcher.py:90-142) -> one evaluation of
    the bundle-adjustment residual + analytic Jacobian (Optimizer.fun, optimizer.py:174-279).  One JSON line with the
    per-stage device times and each stage's roofline; `e2e` = the same job with host descriptors / key points /
    parameters in and host tables / masks / Jacobian out."""
    import torch
    from imageanalysis_b200 import _capi, dist
    rank, world, local = dist.init()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    all_pairs = bates_pairs(args.frames)
    pairs, _, _ = dist.shard_pairs(all_pairs, rank, world)
    P, P_total = len(pairs), len(all_pairs)
    touched = sorted({int(i) for p in pairs for i in p})
    T = len(touched)
    des_u8, uv, ba = make_project_gpu(args.frames, args.desc, SEED, dev, set(touched))
    host = torch.empty((T, args.desc, 128), dtype=torch.float32).pin_memory()
    host_np = host.numpy()
    for k, f in enumerate(touched):
        host_np[k] = des_u8[f]
    eng = _capi.Engine(_capi.NORM_L2, 128, local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    eng.set_profiling(True)
    prm = _capi.Engine.make_params(max_distance=270.0)
    tol = max(1.0, PIPE_W ** 0.25)                           # matcher.py:94-96
    for k, f in enumerate(touched):
        eng.upload(f, host_np[k], pinned=True)
        eng.upload_keypoints(f, uv[f])
    eng.ba_setup(ba["n_cam"], ba["n_pts"], ba["cam_idx"], ba["pt_idx"], ba["uv"])
    eng.ba_upload_params(ba["params"])
    eng.synchronize()
    n_obs = len(ba["cam_idx"])

    def step(evs=None):
        if evs:
            evs[0].record(stream)
        eng.match_pairs_device(pairs, prm)
        if evs:
            evs[1].record(stream)
        eng.ransac_tables(_capi.MODEL_ESSENTIAL, PIPE_K, tol, P, prm.cap, min_pairs=25, compact=True, want_model=False,
                          want_mask=False, host_outputs=False)
        if evs:
            evs[2].record(stream)
        eng.ba_eval_device(ba["K4"], ba["dist"], jac=True)
        if evs:
            evs[3].record(stream)

    for _ in range(max(1, args.warmup)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    launches0 = eng.timing().total_launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    stage_ev = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        step(evs)
        stage_ev.append(evs)
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    t_ms = torch.tensor([ms], device=dev)
    if world > 1:
        torch.distributed.all_reduce(t_ms, op=torch.distributed.ReduceOp.MAX)
    ms = float(t_ms.item())
    tm = eng.timing()
    launches = tm.total_launches - launches0
    st = [statistics.median(e[i].elapsed_time(e[i + 1]) for e in stage_ev) for i in range(3)]
    # results of the last step: inlier statistics of the RANSAC stage, residual of the BA stage
    table, count = eng.fetch_tables(P, prm.cap)
    res = None
    # ---- e2e: host buffers in, host results out ---------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        ids = np.asarray(touched, dtype=np.int32)
        frames_host = [host_np[k] for k in range(T)]
        out_table = torch.empty((P, prm.cap, 2), dtype=torch.int32, pin_memory=True).numpy()
        out_count = torch.zeros((P,), dtype=torch.int32, pin_memory=True).numpy()
        rows_cap = int(count.astype(np.int64).sum()) + 1024       # the filter only removes rows
        out_rows = torch.empty((rows_cap, 2), dtype=torch.int32, pin_memory=True).numpy()
        out_off = torch.empty((P + 1,), dtype=torch.int32, pin_memory=True).numpy()
        out_res = torch.empty((2 * n_obs,), dtype=torch.float64, pin_memory=True).numpy()
        out_jac = torch.empty((n_obs, 2, 10), dtype=torch.float64, pin_memory=True).numpy()

        stage_ms = {}

        def step_e2e():
            t = [time.perf_counter()]
            eng.upload_keypoints_batch(touched, [uv[f] for f in touched])
            t.append(time.perf_counter())
            eng.match_images(ids, frames_host, pairs, prm, out=(out_table, out_count))
            t.append(time.perf_counter())
            _, _, inl = eng.ransac_tables(_capi.MODEL_ESSENTIAL, PIPE_K, tol, P, prm.cap, min_pairs=25, compact=True,
                                          want_model=False)
            t.append(time.perf_counter())
            t2, c2 = eng.fetch_packed_tables(P, out_rows, out_off)    # the filtered match lists, compact (CSR)
            t.append(time.perf_counter())
            r, J = eng.ba_eval(ba["params"], ba["K4"], ba["dist"], jac=True, out=(out_res, out_jac))
            t.append(time.perf_counter())
            for name, a, b in zip(("upload_keypoints", "match_images", "ransac_tables", "fetch_packed_tables", "ba_eval"), t, t[1:]):
                stage_ms[name] = (b - a) * 1e3          # host clock of the last step (every call blocks until its result is on the host)
            return inl, c2, r
        step_e2e()
        torch.cuda.synchronize()
        n_e = max(1, args.steps // 3)
        t0 = time.perf_counter()
        for _ in range(n_e):
            inl, c2, res = step_e2e()
        torch.cuda.synchronize()
        e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev)
        if world > 1:
            torch.distributed.all_reduce(e_ms, op=torch.distributed.ReduceOp.MAX)
        tme = eng.timing()
        h2d = int(tme.h2d_bytes) + T * args.desc * 8 + ba["params"].nbytes
        d2h = P * (prm.cap * 2 + 1) * 4 + int(c2[-1]) * 8 + (P + 1) * 4 + P * 4 + n_obs * (2 + 20) * 8
        e2e = {"value": P_total * n_e / (float(e_ms.item()) / 1e3), "unit": "pairs/s", "ms_per_step": float(e_ms.item()) / n_e,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "host_ms_per_call_last_step": dict(stage_ms),
               "includes": "H2D of float32 descriptors (narrowed on the host), key points and BA parameters; D2H of the match "
                           "tables (padded, as iam_match_images returns them), of the RANSAC-filtered match lists in compact form, "
                           "inlier counts, BA residual and Jacobian blocks"}
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    rl = make_roofline(args, tm.mma_kind, P, tm.knn_ms, peaks, clocks)
    hbm = peaks.get("hbm_gbs") or 6650.0
    ba_gbs = n_obs * 280 / (st[2] / 1e3) / 1e9
    filled = count[count > 0]
    line = {"metric": "image-pairs matched + RANSAC-filtered/sec, with one BA residual+Jacobian evaluation per pass (5000 SIFT desc/img)",
            "value": P_total * args.steps / (ms / 1e3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8 (s32 accumulate) match / f32+f64 RANSAC / f64 BA", "data": "synthetic",
            "config": {"workload": "BASELINE configs[4]: %d frames (38 x 74 survey grid) x %d SIFT descriptors, %d geotag pairs: "
                                   "match + essential-matrix RANSAC per pair + BA residual/Jacobian over %d observations" % (
                                       args.frames, args.desc, P_total, n_obs),
                       "frames": args.frames, "pairs_total": P_total, "pairs_per_gpu": P, "ba_observations": n_obs,
                       "ba_cameras": ba["n_cam"], "ba_points": ba["n_pts"], "ransac_threshold_px": tol,
                       "l2_hygiene": "inputs larger than L2"},
            "stages": {
                "match": {"ms": st[0], "pairs_per_s": P / (st[0] / 1e3), "roofline": rl},
                "ransac_essential": {"ms": st[1], "pairs_per_s": P / (st[1] / 1e3), "pairs_with_matches": int(len(filled)),
                                     "mean_inliers_of_those": float(filled.mean()) if len(filled) else 0.0,
                                     "bound": "FP64 5-point solves + FP32 Sampson scoring (SURVEY 8d); one warp per pair",
                                     "reference": "cv2.findEssentialMat(p1, p2, K, cv2.RANSAC, threshold=tol) (matcher.py:126)"},
                "ba_residual_jacobian": {"ms": st[2], "observations_per_s": n_obs / (st[2] / 1e3),
                                         "roofline": {"bound": "hbm", "achieved": ba_gbs, "peak": hbm, "unit": "GB/s",
                                                      "frac": ba_gbs / hbm, "algorithmic_bytes_per_observation": 280},
                                         "rms_residual_px": float(np.sqrt(np.mean(res ** 2))) if res is not None else None}},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
    print(json.dumps(line))

# ----------------------------------------------------------------------------- main
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload == "pipeline":
        resolve_workload(args, int(os.environ.get("WORLD_SIZE", "1")))
        run_pipeline(args)
        return
    import torch
    from imageanalysis_b200 import _capi, dist
    rank, world, local = dist.init()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    resolve_workload(args, world)
    norm = _capi.NORM_L2 if args.detector == "SIFT" else _capi.NORM_HAMMING
    nbytes = 128 if args.detector == "SIFT" else 32
    bates = args.workload == "bates"

    if bates:   # one project, the pair list sharded: strong scaling.  A rank needs only the frames its block touches.
        all_pairs = bates_pairs(args.frames)
        pairs, p_begin, p_end = dist.shard_pairs(all_pairs, rank, world)
        P_total = len(all_pairs)
        touched = sorted({int(i) for p in pairs for i in p})
        des_u8 = make_frames_gpu(args.frames, args.desc, args.detector, seed=SEED, device=dev, keep=set(touched))
    else:       # every rank its own strip (replicas; the default on one GPU)
        pairs = pair_list(args.frames, args.pairs)
        P_total = len(pairs) * world
        touched = list(range(args.frames))
        des_u8 = make_frames_gpu(args.frames, args.desc, args.detector, seed=SEED + rank, device=dev)
    P = len(pairs)
    T = len(touched)
    # host buffers exactly as the reference holds them: float32 [N,128] for SIFT (image.py:160-180), uint8 for ORB
    host = torch.empty((T, args.desc, nbytes), dtype=torch.float32 if args.detector == "SIFT" else torch.uint8).pin_memory()
    host_np = host.numpy()
    for k, f in enumerate(touched):
        host_np[k] = des_u8[f]

    eng = _capi.Engine(norm, nbytes, local)
    eng.set_engine({"auto": _capi.ENGINE_AUTO, "umma": _capi.ENGINE_UMMA, "umma_f16": _capi.ENGINE_UMMA_F16,
                    "simt": _capi.ENGINE_SIMT}[args.engine])
    # a real (non-default) stream shared by torch and the library, so that the
    # CUDA events below bracket exactly the stream the kernels are launched on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    eng.set_profiling(True)
    prm = _capi.Engine.make_params(max_distance=270.0 if args.detector == "SIFT" else 64.0)

    for k, f in enumerate(touched):
        eng.upload(f, host_np[k], pinned=True)
    eng.synchronize()

    gather_out = None
    gather_ev = []
    gathered = {}
    if world > 1 and args.gather == "padded":
        gather_out = (torch.empty((P_total, prm.cap, 2), dtype=torch.int32, device=dev),
                      torch.empty((P_total,), dtype=torch.int32, device=dev))

    def gather(dt, dc, timed=False):
        """Every rank ends a step holding all ranks' match tables."""
        if world == 1:
            return
        if timed:
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
        c = dist.as_tensor(dc, (P,), local)
        if args.gather == "padded":
            t = dist.as_tensor(dt, (P, prm.cap, 2), local)
            dist.allgather_tables(t, c, P_total, rank, world, out=gather_out)
        else:
            d_rows, _d_off, total = eng.pack_tables_device()
            rows = dist.as_tensor(d_rows, (max(total, 1), 2), local)[:total]
            r_all, c_all = dist.allgather_packed(rows, c, P_total, rank, world)
            gathered["rows"], gathered["count"] = int(r_all.shape[0]), int(c_all.shape[0])
        if timed:
            g1.record(stream)
            gather_ev.append((g0, g1))

    def step_device(timed=False):
        dt, dc = eng.match_pairs_device(pairs, prm)
        gather(dt, dc, timed)

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    launches0 = eng.timing().total_launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record(stream)
    knn_samples, reduce_samples = [], []
    for _ in range(args.steps):
        step_device(timed=True)
        # the library brackets its kernels with CUDA events on this stream; reading them back waits for the step's
        # last kernel (a few microseconds of launch latency per 16 ms step), so every timed launch is in the average
        t_step = eng.timing()
        knn_samples.append(t_step.knn_ms)
        reduce_samples.append(t_step.reduce_ms)
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    gather_ms = sum(a.elapsed_time(b) for a, b in gather_ev) / max(1, len(gather_ev)) if gather_ev else 0.0
    tm = eng.timing()
    launches = tm.total_launches - launches0
    # average launch duration of the dominant kernel over the timed region
    knn_kernel_ms = sum(knn_samples) / len(knn_samples)
    reduce_ms = sum(reduce_samples) / len(reduce_samples)
    t_ms = torch.tensor([ms], device=dev)
    if world > 1:
        torch.distributed.all_reduce(t_ms, op=torch.distributed.ReduceOp.MAX)
    ms = float(t_ms.item())
    value = P_total * args.steps / (ms / 1e3)

    # ---- untimed: a few tables of this run against the CPU oracle ----------------
    spot = None
    if not args.no_spot and rank == 0:
        eng.match_pairs_device(pairs, prm)
        table, count = eng.fetch_tables(P, prm.cap)
        filled = np.nonzero(count > 0)[0]
        which = sorted({0, int(filled[len(filled) // 2]) if len(filled) else P // 2, P - 1})
        spot = parity_spot(des_u8, pairs, table, count, args.detector, which)
        spot["mean_matches_per_pair"] = float(count.mean())
        del table

    # ---- end to end through the public API with host buffers -----------------
    e2e = None
    if not args.no_e2e:
        h2d = int(host_np.nbytes)
        d2h = int(P * (prm.cap * 2 + 1) * 4)
        ids = np.asarray(touched, dtype=np.int32)
        frames_host = [host_np[k] for k in range(T)]
        # result buffers a pipeline would keep around: page-locked, so each wave's tables come back asynchronously
        out_table = torch.empty((P, prm.cap, 2), dtype=torch.int32, pin_memory=True).numpy()
        out_count = torch.zeros((P,), dtype=torch.int32, pin_memory=True).numpy()

        # The call uploads a frame right before the first pair that needs it, so the ORDER of the pair list decides
        # how soon matching can start: sorted by the later frame of each pair, a pair is ready as soon as that frame
        # has crossed the bus (the (i, j)-sorted list needs frames two flight lines ahead for its very first pairs:
        # 45 % of the upload would pass before the first kernel).  Tables come back in the order given.
        pairs_e2e = np.ascontiguousarray(pairs[np.lexsort((pairs[:, 0], pairs[:, 1]))])

        def step_e2e():
            # the public one-call API: host descriptors in, host match tables out (H2D + conversion + matching + D2H);
            # several GPUs: plus the gather, so that every rank ends the step holding every table on its device
            r = eng.match_images(ids, frames_host, pairs_e2e, prm, out=(out_table, out_count))
            if world > 1:
                d_rows, _d_off, total = eng.pack_tables_device()
                rows = dist.as_tensor(d_rows, (max(total, 1), 2), local)[:total]
                c = torch.from_numpy(out_count).to(dev, non_blocking=True)
                dist.allgather_packed(rows, c, P_total, rank, world)
            return r

        step_e2e()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        n_e = max(1, args.steps // 2)
        t0 = time.perf_counter()
        for _ in range(n_e):
            table, count = step_e2e()
        torch.cuda.synchronize()
        e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev)
        if world > 1:
            torch.distributed.all_reduce(e_ms, op=torch.distributed.ReduceOp.MAX)
        tme = eng.timing()
        # bytes that actually crossed PCIe in the last step: float32 frames the library's worker threads narrowed to
        # uint8 on the host (integer-valued SIFT descriptors, transport only) count as bytes
        bytes_t = torch.tensor([float(int(tme.h2d_bytes) or h2d), float(d2h), float(h2d), float(tme.narrowed_images), float(T)],
                               device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(bytes_t)
        # context: the same call fed with uint8 descriptors -- what this framework's own detector hands out
        # (detector.SIFT_create(uint8_descriptors=True)); the headline e2e above keeps the reference's float32 arrays
        u8_leg = None
        if args.detector == "SIFT":
            host_u8 = torch.empty((T, args.desc, 128), dtype=torch.uint8).pin_memory()
            host_u8.numpy()[:] = host_np                    # integer-valued: exact
            frames_u8 = [host_u8.numpy()[k] for k in range(T)]
            sample = range(0, P, max(1, P // 64))
            ref_count = np.array(count)                     # the float32 run's results (the output arrays are reused)
            ref_rows = {p: np.array(table[p, :count[p]]) for p in sample}
            frames_f32, frames_host = frames_host, frames_u8
            step_e2e()
            torch.cuda.synchronize()
            if world > 1:
                torch.distributed.barrier()
            t0 = time.perf_counter()
            for _ in range(n_e):
                table_u8, count_u8 = step_e2e()
            torch.cuda.synchronize()
            u_ms = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev)
            if world > 1:
                torch.distributed.all_reduce(u_ms, op=torch.distributed.ReduceOp.MAX)
            frames_host = frames_f32
            same = bool(np.array_equal(count_u8, ref_count) and all(
                np.array_equal(table_u8[p, :ref_count[p]], ref_rows[p]) for p in sample))
            u8_leg = {"value": P_total * n_e / (float(u_ms.item()) / 1e3), "unit": "pairs/s", "ms_per_step": float(u_ms.item()) / n_e,
                      "h2d_bytes_per_step_this_rank": int(eng.timing().h2d_bytes), "tables_equal_float32_run": same,
                      "timeline_ms_rank0": {"host_enqueue": eng.timing().host_enqueue_ms, "upload_span": eng.timing().upload_span_ms,
                                            "compute_span": eng.timing().compute_span_ms},
                      "note": "context only: uint8 host descriptors as detector.SIFT_create(uint8_descriptors=True) returns them"}
            del host_u8, frames_u8
        # context for the e2e number: what a bare pinned-host -> device copy of the same bytes costs on this box
        dst = torch.empty_like(host, device=dev)
        dst.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        c0 = time.perf_counter()
        dst.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        h2d_ms = (time.perf_counter() - c0) * 1e3
        del dst
        e2e = {"value": P_total * n_e / (float(e_ms.item()) / 1e3), "unit": "pairs/s",
               "h2d_bytes_per_step": int(bytes_t[0].item()), "d2h_bytes_per_step": int(bytes_t[1].item()),
               "ms_per_step": float(e_ms.item()) / n_e,
               "host_input_bytes_per_step": int(bytes_t[2].item()), "frames_narrowed_on_host": int(bytes_t[3].item()),
               "frames_uploaded_per_step": int(bytes_t[4].item()),
               "host_threads": int(os.environ.get("IAM_HOST_THREADS", "0")) or None,
               "mean_matches_per_pair": float(count.mean()), "bare_h2d_ms_same_bytes_rank0": h2d_ms,
               "timeline_ms_rank0": {"host_enqueue": tme.host_enqueue_ms, "upload_span": tme.upload_span_ms,
                                     "compute_span": tme.compute_span_ms, "total_span": tme.total_span_ms,
                                     "waves": tme.waves},
               "bare_h2d_gb_per_s_rank0": h2d / h2d_ms / 1e6, "uint8_input": u8_leg,
               "pair_order": "this rank's block sorted by the later frame of each pair (upload-friendly; tables returned in that order)",
               "includes": "H2D of every touched frame's float32 descriptors + conversion + matching + D2H of this rank's "
                           "tables" + (" + compact all-gather of all tables" if world > 1 else "")}

    # ---- BASELINE configs[2]: a short ORB leg folded into the default line ---------
    orb = None
    if not args.no_orb and world == 1 and args.detector == "SIFT" and args.workload == "strip":
        orb = orb_leg(args, dev, local, stream, tm.mma_kind)

    # ---- the detect stage in front of the matcher (image.py:236-237, :324): a short SIFT leg ----
    sift = None
    if not args.no_orb and world == 1 and args.detector == "SIFT" and args.workload == "strip":
        sift = sift_leg(local)

    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    roofline = make_roofline(args, tm.mma_kind, P, knn_kernel_ms, peaks, clocks)
    roofline.update({"reduce_ms_per_step": reduce_ms, "allgather_ms_per_step": gather_ms,
                     "kernel_ms_min_max": [min(knn_samples), max(knn_samples)],
                     "engine": {1: "umma", 2: "simt"}.get(tm.engine_used),
                     "gather": None if world == 1 else dict(gathered, mode=args.gather)})
    cpu = None
    if not args.no_cpu and world == 1:
        cpu = cpu_pairs_per_s(des_u8, pairs, args.detector, args.cpu_pairs, repeats=2, context=True)
    line = {"metric": "image-pairs matched/sec (5000 %s desc/img)" % args.detector, "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if bates else "weak", "vs_baseline": None,
            "dtype": {0: "f16", 1: "e4m3", 2: "u8 (s32 accumulate)"}.get(tm.mma_kind, "u8"), "data": "synthetic",
            "config": workload_config(args, P_total, world, P, T), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "parity_spot": spot, "orb": orb, "sift_detect": sift, "gpu_launches": launches, "clocks": clocks}
    print(json.dumps(line))


def make_roofline(args, mma_kind_id, n_pairs, knn_kernel_ms, peaks, clocks, detector=None):
    """Tensor roofline of the kNN kernel: algorithmic work (one N x M x D product per pair) over the kernel's
    own time (CUDA events on the launching stream), against the measured peak of the MMA kind the kernel issues."""
    detector = detector or args.detector
    capped = bool(clocks and "sw_power_cap" in (clocks.get("reasons") or []))
    burst, sustained = peaks.get("bf16_tflops"), peaks.get("bf16_tflops_sustained")
    if peaks:
        bf16_peak = (sustained if capped and sustained else burst) or 1590.0
        src = "measured (MEASURED_PEAKS.json %s: %s)" % (
            "bf16_tflops_sustained" if capped and sustained else "bf16_tflops",
            "sw_power_cap was sampled during the timed region" if capped else
            "no power cap sampled during the timed region, SM clock at %s MHz" % (clocks or {}).get("sm_mhz"))
    else:
        bf16_peak, src = 1590.0, "fallback (B200_PROFILING.md burst 1.59 PFLOP/s)"
    work = FLOP_PER_PAIR_L2 if detector == "SIFT" else OP_PER_PAIR_HAMMING
    work *= (args.desc / 5000.0) ** 2
    achieved = n_pairs * work / (knn_kernel_ms / 1e3) / 1e12 if knn_kernel_ms > 0 else None
    traffic = None
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tj) and detector == "SIFT":
        traffic = json.load(open(tj)).get("knn_umma_dram_bytes_per_launch")
    mma_kind = {0: "kind::f16", 1: "kind::f8f6f4", 2: "kind::i8"}.get(mma_kind_id, "none (SIMT)")
    # kind::i8 and kind::f8f6f4 issue at twice the f16 rate (K = 32 bytes per instruction in the same cycles; nominal
    # 4.5 vs 2.25 P dense).  There is no measured int8 / fp8 GEMM peak in MEASURED_PEAKS.json: the kind's peak is taken
    # as 2 x the measured bf16 peak, and the bf16 figures are kept next to it as context.
    kind_rate = 2.0 if mma_kind_id in (1, 2) else 1.0
    peak = bf16_peak * kind_rate
    return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s" if mma_kind_id == 0 else "TOP/s",
            "frac": (achieved / peak) if achieved else None, "traffic": traffic,
            "peak_source": src + ("; x2 for %s" % mma_kind if kind_rate == 2.0 else ""),
            "kernel": "knn_umma_kernel (tcgen05 %s)" % mma_kind, "mma_kind": mma_kind, "kind_rate_vs_bf16": kind_rate,
            "context_bf16": {"peak_burst": burst, "peak_sustained": sustained,
                             "frac_of_bf16_burst": (achieved / burst) if achieved and burst else None,
                             "frac_of_bf16_sustained": (achieved / sustained) if achieved and sustained else None},
            "kernel_ms_per_launch": knn_kernel_ms, "algorithmic_work_per_pair": work,
            "issued_work_note": "the kernel issues one product per DIRECTION (2 per pair) with K = 160 (L2 bytes) / 288 "
                                "(Hamming e4m3) instead of 128 / 256: tensor-pipe activity is ~2.6x / 2.3x the algorithmic share"}


def orb_leg(args, dev, local, stream, _kind):
    """BASELINE configs[2]: 500 frames x 5000 ORB descriptors (256 bit), 1990 pairs, device-resident, 5 timed steps;
    one pair against the CPU oracle.  Reported inside the default line as "orb"."""
    import torch
    from imageanalysis_b200 import _capi
    frames, n = 500, args.desc
    des = make_frames_gpu(frames, n, "ORB", seed=SEED, device=dev)
    pairs = pair_list(frames, "sequential")
    eng = _capi.Engine(_capi.NORM_HAMMING, 32, local)
    eng.set_stream(stream.cuda_stream)
    eng.set_profiling(True)
    prm = _capi.Engine.make_params(max_distance=64.0)
    for f in range(frames):
        eng.upload(f, des[f])
    eng.synchronize()
    for _ in range(2):
        eng.match_pairs_device(pairs, prm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 5
    e0.record(stream)
    for _ in range(steps):
        eng.match_pairs_device(pairs, prm)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    tm = eng.timing()
    table, count = eng.fetch_tables(len(pairs), prm.cap)
    spot = parity_spot(des, pairs, table, count, "ORB", [0, len(pairs) - 1])
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    rl = make_roofline(args, tm.mma_kind, len(pairs), tm.knn_ms, peaks, None, detector="ORB")
    eng.close()
    return {"config": "BASELINE configs[2]: 500 frames x %d ORB descriptors (256 bit), Hamming, 1990 pairs, device-resident" % n,
            "value": len(pairs) / (ms / 1e3), "unit": "pairs/s", "ms_per_step": ms, "steps": steps,
            "kernel_ms_per_launch": tm.knn_ms, "mma_kind": rl["mma_kind"], "achieved_tops": rl["achieved"],
            "frac": rl["frac"], "frac_of_bf16_burst": rl["context_bf16"]["frac_of_bf16_burst"],
            "parity_spot": spot["status"], "mean_matches_per_pair": float(count.mean())}


def survey_frame(h=1459, w=2189, seed=3):
    """A synthetic grey frame at the reference's detection size (0.4 x a 5472 x 3648 photo): band-limited noise."""
    rng = np.random.default_rng(seed)
    f = rng.integers(0, 256, (h, w)).astype(np.float32)
    k = np.exp(-0.5 * (np.arange(-6, 7) / 2.0) ** 2).astype(np.float32)
    k /= k.sum()
    for axis in (0, 1):
        f = np.apply_along_axis(lambda r: np.convolve(r, k, "same"), axis, f)
    return ((f - f.min()) / (f.max() - f.min()) * 255).astype(np.uint8)


def sift_leg(local):
    """The detector in front of the matcher: cv2.SIFT_create().detectAndCompute of one survey-sized frame through
    detector.SIFT's C-ABI call (iam_sift_detect; host image in, host key points + descriptors out), with cv2 timed beside
    it on the host cores and the key points / descriptors compared when cv2 is importable.  Reported as "sift_detect"."""
    from imageanalysis_b200 import _capi
    img = survey_frame()
    eng = _capi.Engine(_capi.NORM_L2, 128, local)
    for _ in range(3):
        kp, octv, des = eng.sift_detect(img)
    l0 = eng.timing().total_launches
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        kp, octv, des = eng.sift_detect(img)
    ms = (time.perf_counter() - t0) * 1e3 / reps
    launches = (eng.timing().total_launches - l0) // reps
    eng.close()
    # several frames in flight (one context, host thread and stream each): what a project's detect pass sees
    from imageanalysis_b200 import detector
    frames = [img] * 12
    detector.sift_detect_many(frames[:4], workers=2)
    t0 = time.perf_counter()
    detector.sift_detect_many(frames, workers=2)
    ms_many = (time.perf_counter() - t0) * 1e3 / len(frames)
    out = {"config": "one 2189 x 1459 grey frame (the reference's 0.4 scale of a 5472 x 3648 photo), OpenCV's default SIFT "
                     "parameters, host image in -> host key points + uint8 descriptors out",
           "ms_per_frame": ms, "frames_per_s": 1e3 / ms, "keypoints": int(len(kp)), "gpu_launches_per_frame": int(launches),
           "ms_per_frame_two_in_flight": ms_many, "frames_per_s_two_in_flight": 1e3 / ms_many,
           "timing": "host clock around the blocking C-ABI call, mean of %d" % reps}
    try:
        import cv2
    except ImportError:
        return out
    det = cv2.SIFT_create()
    t0 = time.perf_counter()
    k2, d2 = det.detectAndCompute(img, None)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    from oracle import sift as osift     # the checker, outside every timed region
    kr = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response] for k in k2], np.float32).reshape(-1, 5)
    m = osift.match_keypoints(kr, kp)
    ok = m >= 0
    dd = np.abs(d2[ok].astype(np.int32) - des[m[ok]].astype(np.int32)).max(axis=1) if ok.any() else np.zeros(0, np.int32)
    out["cpu_baseline"] = {"value": 1e3 / cpu_ms, "unit": "frames/s", "cores": cv2.getNumThreads(), "kind": "reference",
                           "sample": "cv2.SIFT_create().detectAndCompute on the same frame, one call (%.0f ms)" % cpu_ms}
    out["parity_vs_cv2"] = {"keypoints_cv2": int(len(kr)), "reproduced_within_0.02px_0.5deg": int(ok.sum()),
                            "descriptors_identical": int((dd == 0).sum()), "descriptors_within_1": int((dd <= 1).sum()),
                            "tolerance": "float pipeline: see tests/test_sift.py"}
    return out


if __name__ == "__main__":
    main()
