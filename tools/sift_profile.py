#!/usr/bin/env python3
"""Run iam_sift_detect a few times on one survey-sized synthetic image (for an ncu launch list)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from imageanalysis_b200 import detector  # noqa: E402

h, w = (int(sys.argv[2]), int(sys.argv[1])) if len(sys.argv) > 2 else (1459, 2189)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rng = np.random.default_rng(5)
f = rng.integers(0, 256, (h, w)).astype(np.float32)
k = np.exp(-0.5 * (np.arange(-8, 9) / 2.5) ** 2)
k /= k.sum()
f = np.apply_along_axis(lambda r: np.convolve(r, k, "same"), 1, f)
f = np.apply_along_axis(lambda r: np.convolve(r, k, "same"), 0, f)
img = ((f - f.min()) / (f.max() - f.min()) * 255).astype(np.uint8)
eng = detector._eng()
for it in range(reps):
    t0 = time.perf_counter()
    kp, octv, des = eng.sift_detect(img)
    print("%dx%d: %d key points, %.2f ms" % (w, h, len(kp), 1e3 * (time.perf_counter() - t0)), flush=True)
