#!/usr/bin/env python3
"""bench.py — image-pairs matched per second on BASELINE.json's config.

A "step" is one pass of the hot path over one batch: every candidate pair of a
500-frame strip with 5000 SIFT descriptors per frame (BASELINE configs[1];
pair list = the reference's live generator, |i-j| <= 4 -> 1990 pairs; use
`--pairs all` for the 124 750-pair upper triangle) goes through
kNN (both directions) -> metric reduction -> cross-check -> per-pair match table.

  value : pairs/s with descriptors already resident in HBM (device timed, CUDA events)
  e2e   : the same job through the public API with HOST buffers: H2D of every
          frame's float32 descriptors from pinned memory + layout conversion +
          matching + D2H of the match tables, every step
  roofline : algorithmic 2*N*M*128 FLOP per pair / the kNN kernel's own time
  cpu_baseline / --impl reference : the reference's CPU path (cv2.BFMatcher
          both directions + the Python reduction of matcher.py:253-269) on the
          box's host cores, on a bounded sample of the same pairs

Multi-GPU (torchrun, one rank per GPU): every rank matches its own 500-frame
strip (weak scaling) and one NCCL all-gather collects all match tables.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=500)
    ap.add_argument("--desc", type=int, default=5000)
    ap.add_argument("--detector", default="SIFT", choices=["SIFT", "ORB"])
    ap.add_argument("--pairs", default="sequential", choices=["sequential", "all"])
    ap.add_argument("--workload", default="strip", choices=["strip", "bates"],
                    help="strip: BASELINE configs[1] per GPU (weak scaling, the default); bates: configs[3], 2812 frames "
                         "on a 38 x 74 serpentine survey grid, geotag-neighbour pair list sharded across the ranks")
    ap.add_argument("--engine", default="auto", choices=["auto", "umma", "umma_f16", "simt"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-pairs", type=int, default=4)
    return ap.parse_args()


# ----------------------------------------------------------------------------- data
def make_frames_gpu(frames, n, detector, seed, device):
    """Synthetic descriptors with OpenCV-SIFT statistics (see imageanalysis_b200/synth.py),
    generated with torch for speed, returned as a HOST uint8 array [frames, n, D]."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = []
    prev = None
    for f in range(frames):
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
        out.append(d.cpu())
    return torch.stack(out).numpy()


def pair_list(frames, mode):
    if mode == "all":
        return np.asarray([(i, j) for i in range(frames) for j in range(i + 1, frames)], np.int32)
    return np.asarray([(i, j) for i in range(frames) for j in range(i + 1, min(frames, i + 5))], np.int32)


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
    """The reference's CPU path for `n_pairs` pairs: knnMatch both ways with the
    exact matcher (cv2.BFMatcher when cv2 is importable: kind 'reference';
    otherwise the C restatement oracle/oracle_knn.c: kind 'port') plus the
    Python reduction of matcher.py:253-269 and the cross-check (:187-200)."""
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
           "sample": "%d of the %d pairs of this workload, both kNN directions + matcher.py:253-269 reduction + "
                     "cross-check, best of %d" % (len(sample), len(pairs), repeats)}
    if context and kind == "reference":
        # Context only: the same pair on ONE host thread, and the matcher the reference literally
        # configures -- an approximate FLANN kd-tree search (matcher.py:62-79), kNN both directions without the reduction.
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
                fl = cv2.FlannBasedMatcher(dict(algorithm=1, trees=5), dict(checks=50))
                t0 = time.perf_counter()
                fl.knnMatch(a, b, k=2)
                fl.knnMatch(b, a, k=2)
                ctx["flann_kdtree_knn_pairs_per_s"] = 1.0 / (time.perf_counter() - t0)
            out["context"] = ctx
        except Exception as e:  # noqa: BLE001  (context figures must never break the bench line)
            out["context"] = {"error": str(e)[:80]}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from imageanalysis_b200 import synth
    frames = min(args.frames, args.cpu_pairs // 4 + 6)
    gen = synth.sift_like if args.detector == "SIFT" else synth.orb_like
    rng = np.random.default_rng(0)
    des = []
    for f in range(frames):
        d = gen(args.desc, seed=1000 + f)
        if f > 0:
            synth.plant(des[-1], d, 0.4, rng, "sift" if args.detector == "SIFT" else "orb")
        des.append(d)
    pairs = pair_list(frames, "sequential")
    steps = []
    base = None
    for s in range(args.warmup + args.steps):
        base = cpu_pairs_per_s(des, pairs, args.detector, args.cpu_pairs, repeats=0 if s else 1,
                               context=(s == args.warmup + args.steps - 1))
        if s >= args.warmup:
            steps.append(base["value"])
    v = statistics.median(steps)
    base["value"] = v
    line = {"impl": "reference", "metric": "image-pairs matched/sec (5000 %s desc/img)" % args.detector, "value": v,
            "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * args.cpu_pairs / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.detector == "SIFT" else "u8", "data": "synthetic",
            "config": workload_config(args, len(pair_list(args.frames, args.pairs)), 1),
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def bates_pairs(frames):
    """BASELINE configs[3]: 38 flight lines x 74 frames, 15 m along-track, 25 m cross-track,
    serpentine; pairs = the reference's camera-distance window (matcher.py:858-903, pairs.worklist 'geotag')."""
    from imageanalysis_b200 import pairs as wl, synth
    neds = synth.survey_grid_neds()[:frames]
    return wl.pair_array(wl.worklist(neds, "geotag"))


def workload_config(args, n_pairs, world):
    if args.workload == "bates":
        return {"workload": "%d frames (38 x 74 serpentine survey grid) x %d %s descriptors/frame, geotag-neighbour pair "
                            "list (reference matcher.py:858-903), %d pairs in all, sharded across %d GPU(s)" % (
                                args.frames, args.desc, args.detector, n_pairs, world),
                "frames": args.frames, "desc_per_frame": args.desc, "pairs_total": n_pairs,
                "pairs_per_gpu": -(-n_pairs // world), "match_ratio": 0.75, "cap": 2000, "min_pairs": 25,
                "l2_hygiene": "inputs larger than L2 (byte forms %.2f GB per GPU, replicated)" % (
                    args.frames * (-(-args.desc // 256) * 256) * 192 / 1e9),
                "parallelism": "pair-sharded x%d + 1 NCCL all-gather of match tables" % world if world > 1 else "single GPU"}
    return {"workload": "%d frames x %d %s descriptors/frame, %s pair list (%d pairs) per GPU" % (
                args.frames, args.desc, args.detector,
                "|i-j|<=4 (reference matcher.py:899)" if args.pairs == "sequential" else "all-pairs", n_pairs),
            "frames_per_gpu": args.frames, "desc_per_frame": args.desc, "pairs_per_gpu": n_pairs,
            "pairs_total": n_pairs * world, "match_ratio": 0.75, "cap": 2000, "min_pairs": 25,
            "l2_hygiene": "inputs larger than L2 (operand forms %.2f GB per GPU)" % (
                args.frames * 2 * (-(-args.desc // 256) * 256) * 288 / 1e9),
            "parallelism": "pair-sharded x%d + 1 NCCL all-gather of match tables" % world if world > 1 else "single GPU"}


# ----------------------------------------------------------------------------- main
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    from imageanalysis_b200 import _capi, dist
    rank, world, local = dist.init()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    norm = _capi.NORM_L2 if args.detector == "SIFT" else _capi.NORM_HAMMING
    nbytes = 128 if args.detector == "SIFT" else 32

    bates = args.workload == "bates"
    if bates:  # every rank holds every frame (replicated descriptors), the pair list is sharded: strong scaling
        if args.frames == 500:
            args.frames = 2812
        args.no_e2e = args.no_cpu = True
        des_u8 = make_frames_gpu(args.frames, args.desc, args.detector, seed=1234, device=dev)
        host_np = des_u8
        all_pairs = bates_pairs(args.frames)
        pairs, _, _ = dist.shard_pairs(all_pairs, rank, world)
        P_total = len(all_pairs)
    else:
        des_u8 = make_frames_gpu(args.frames, args.desc, args.detector, seed=1234 + rank, device=dev)
        # host buffers exactly as the reference holds them: float32 [N,128] for SIFT (image.py:160-180), uint8 for ORB
        host = torch.from_numpy(des_u8.astype(np.float32) if args.detector == "SIFT" else des_u8).pin_memory()
        host_np = host.numpy()
        pairs = pair_list(args.frames, args.pairs)
        P_total = len(pairs) * world
    P = len(pairs)

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

    def upload_all():
        for f in range(args.frames):
            eng.upload(f, host_np[f], pinned=True)

    gather_out = None
    gather_ev = []
    if world > 1:  # every rank ends a step holding all ranks' tables
        gather_out = (torch.empty((P_total, prm.cap, 2), dtype=torch.int32, device=dev),
                      torch.empty((P_total,), dtype=torch.int32, device=dev))

    def step_device(timed=False):
        dt, dc = eng.match_pairs_device(pairs, prm)
        if world > 1:
            t = dist.as_tensor(dt, (P, prm.cap, 2), local)
            c = dist.as_tensor(dc, (P,), local)
            if timed:
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record(stream)
            dist.allgather_tables(t, c, P_total, rank, world, out=gather_out)
            if timed:
                g1.record(stream)
                gather_ev.append((g0, g1))

    upload_all()
    eng.synchronize()
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
    knn_ms = []
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(args.steps):
        step_device(timed=True)
        knn_ms.append(None)
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    gather_ms = sum(a.elapsed_time(b) for a, b in gather_ev) / max(1, len(gather_ev)) if gather_ev else 0.0
    tm = eng.timing()
    launches = tm.total_launches - launches0
    # per-kernel time of the dominant kernel (events recorded by the library on the same stream)
    knn_kernel_ms = tm.knn_ms
    reduce_ms = tm.reduce_ms
    t_ms = torch.tensor([ms], device=dev)
    if world > 1:
        torch.distributed.all_reduce(t_ms, op=torch.distributed.ReduceOp.MAX)
    ms = float(t_ms.item())
    value = P_total * args.steps / (ms / 1e3)

    # ---- end to end through the public API with host buffers -----------------
    e2e = None
    if not args.no_e2e:
        h2d = int(host_np.nbytes)
        d2h = int(P * (prm.cap * 2 + 1) * 4)

        ids = np.arange(args.frames, dtype=np.int32)
        frames_host = [host_np[f] for f in range(args.frames)]

        # result buffers a pipeline would keep around: page-locked, so each wave's tables come back asynchronously
        out_table = torch.empty((P, prm.cap, 2), dtype=torch.int32, pin_memory=True).numpy()
        out_count = torch.zeros((P,), dtype=torch.int32, pin_memory=True).numpy()

        def step_e2e():
            # the public one-call API: host descriptors in, host match tables out (H2D + conversion + matching + D2H)
            return eng.match_images(ids, frames_host, pairs, prm, out=(out_table, out_count))
        step_e2e()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        for _ in range(max(1, args.steps // 2)):
            table, count = step_e2e()
        torch.cuda.synchronize()
        e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev)
        if world > 1:
            torch.distributed.all_reduce(e_ms, op=torch.distributed.ReduceOp.MAX)
        n_e = max(1, args.steps // 2)
        # context for the e2e number: what a bare pinned-host -> device copy of the same bytes costs on this box
        dst = torch.empty_like(host, device=dev)
        dst.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        c0 = time.perf_counter()
        dst.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        h2d_ms = (time.perf_counter() - c0) * 1e3
        del dst
        # bytes that actually crossed PCIe in the last step: float32 frames the library's worker threads narrowed to
        # uint8 on the host (integer-valued SIFT descriptors, transport only) count as bytes
        tme = eng.timing()
        e2e = {"value": P * world * n_e / (float(e_ms.item()) / 1e3), "unit": "pairs/s",
               "h2d_bytes_per_step": int(tme.h2d_bytes) or h2d, "d2h_bytes_per_step": d2h, "ms_per_step": float(e_ms.item()) / n_e,
               "host_input_bytes_per_step": h2d, "frames_narrowed_on_host": int(tme.narrowed_images),
               "host_threads": int(os.environ.get("IAM_HOST_THREADS", "0")) or None,
               "mean_matches_per_pair": float(count.mean()), "bare_h2d_ms_same_bytes": h2d_ms,
               "timeline_ms": {"host_enqueue": eng.timing().host_enqueue_ms, "upload_span": eng.timing().upload_span_ms,
                               "compute_span": eng.timing().compute_span_ms, "total_span": eng.timing().total_span_ms,
                               "waves": eng.timing().waves},
               "bare_h2d_gb_per_s": h2d / h2d_ms / 1e6}

    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained; kernel timed inside a %.0f ms step)" % (ms / args.steps) \
        if peaks else "fallback (B200_PROFILING.md sustained 1.4 PFLOP/s)"
    work = FLOP_PER_PAIR_L2 if args.detector == "SIFT" else OP_PER_PAIR_HAMMING
    work *= (args.desc / 5000.0) ** 2
    achieved = P * work / (knn_kernel_ms / 1e3) / 1e12 if knn_kernel_ms > 0 else None  # rank 0's shard, rank 0's kernel time
    traffic = None
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tj):
        traffic = json.load(open(tj)).get("knn_umma_dram_bytes_per_launch")
    mma_kind = {0: "kind::f16", 1: "kind::f8f6f4", 2: "kind::i8"}.get(tm.mma_kind, "none (SIMT)")
    kind_rate = 2.0 if tm.mma_kind in (1, 2) else 1.0
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "peak_burst": peaks.get("bf16_tflops"),
                "frac_of_burst": (achieved / peaks["bf16_tflops"]) if achieved and peaks.get("bf16_tflops") else None,
                "kernel": "knn_umma_kernel (tcgen05 %s)" % mma_kind,
                # The contract's denominator is the measured dense bf16 peak.  kind::i8 / kind::f8f6f4 issue at twice the
                # f16 rate (nominal 4.5 vs 2.25 P dense): the fraction of THAT pipe's peak is stated next to it.
                "mma_kind": mma_kind, "kind_rate_vs_bf16": kind_rate,
                "frac_of_kind_peak": (achieved / (peak * kind_rate)) if achieved else None,
                "kernel_ms_per_launch": knn_kernel_ms, "reduce_ms_per_step": reduce_ms,
                "algorithmic_work_per_pair": work, "allgather_ms_per_step": gather_ms,
                "engine": {1: "umma", 2: "simt"}.get(tm.engine_used)}
    cpu = None
    if not args.no_cpu and world == 1:
        cpu = cpu_pairs_per_s(des_u8, pairs, args.detector, args.cpu_pairs, repeats=2, context=True)
    line = {"metric": "image-pairs matched/sec (5000 %s desc/img)" % args.detector, "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if bates else "weak", "vs_baseline": None,
            "dtype": {0: "f16", 1: "e4m3", 2: "u8 (s32 accumulate)"}.get(tm.mma_kind, "u8"), "data": "synthetic",
            "config": workload_config(args, P_total if bates else P, world), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
