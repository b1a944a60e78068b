#!/usr/bin/env python3
"""Debug aid: GPU SIFT (iam_sift_detect) against the cv2 goldens and the CPU restatement, plus timing at survey sizes."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from imageanalysis_b200 import detector  # noqa: E402
from oracle import sift as S  # noqa: E402

g = np.load("tests/golden/sift_reference.npz")


def report(tag, kp, octv, des, kr, orf, dr):
    m = S.match_keypoints(kr, kp)
    ok = m >= 0
    dd = np.abs(dr[ok].astype(int) - des[m[ok]].astype(int)).max(axis=1) if ok.any() else np.zeros(0, int)
    print("%s: ref %d ours %d reproduced %d (%.2f%%) octave eq %s  des identical %d  <=1 %d  <=2 %d  worst %d" % (
        tag, len(kr), len(kp), ok.sum(), 100.0 * ok.mean(), np.array_equal(orf[ok] & 0xFFFF, octv[m[ok]] & 0xFFFF),
        (dd == 0).sum(), (dd <= 1).sum(), (dd <= 2).sum(), dd.max() if len(dd) else -1), flush=True)
    miss = np.nonzero(~ok)[0][:5]
    for i in miss:
        d = np.abs(kp[:, 0] - kr[i, 0]) + np.abs(kp[:, 1] - kr[i, 1])
        j = int(np.argmin(d))
        print("   missing", kr[i], hex(int(orf[i])), " nearest ours", kp[j], hex(int(octv[j])))


for name in ("small", "medium", "large"):
    img = g[name + "_image"]
    r = detector.sift_detect_and_compute(img)
    kp = np.column_stack([r["pt"], r["size"], r["angle"], r["response"]]).astype(np.float32)
    report(name + " vs cv2", kp, r["octave"], r["des"], g[name + "_kp"], g[name + "_octave"], g[name + "_des"])
    if name == "small":
        ko, oo, do = S.detect_arrays(img)
        report(name + " vs restatement", kp, r["octave"], r["des"], ko, oo, do)

try:
    import cv2
except ImportError:
    cv2 = None
rng = np.random.default_rng(5)
for (h, w) in ((729, 1094), (1459, 2189)):
    f = rng.integers(0, 256, (h, w)).astype(np.float32)
    if cv2 is not None:
        f = cv2.GaussianBlur(f, (0, 0), 2.5)
    img = ((f - f.min()) / (f.max() - f.min()) * 255).astype(np.uint8)
    eng = detector._eng()
    for it in range(4):
        l0 = eng.timing().total_launches
        t0 = time.perf_counter()
        kp, octv, des = eng.sift_detect(img)
        t1 = time.perf_counter()
        print("%dx%d: %d key points, %.2f ms host to host, %d launches" % (w, h, len(kp), 1e3 * (t1 - t0), eng.timing().total_launches - l0), flush=True)
    if cv2 is not None:
        s = cv2.SIFT_create()
        t0 = time.perf_counter()
        k2, d2 = s.detectAndCompute(img, None)
        t1 = time.perf_counter()
        print("   cv2.SIFT_create().detectAndCompute: %d key points, %.1f ms (%d threads)" % (len(k2), 1e3 * (t1 - t0), cv2.getNumThreads()))
        kr = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response] for k in k2], np.float32)
        report("   %dx%d vs live cv2" % (w, h), kp, octv, des, kr, np.array([k.octave for k in k2], np.int32), d2.astype(np.uint8))
