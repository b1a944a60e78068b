#!/usr/bin/env python3
"""Golden fixtures for the SIFT detect + describe stage (reference scripts/lib/image.py:236-237, :324:
`cv2.SIFT_create().detectAndCompute(scaled, None)`), recorded from LIVE cv2 in the build container.

sift_reference.npz -- three synthetic grey images (band-limited noise, the same generator the ORB goldens use, so
that key points appear at every octave) and what cv2.SIFT_create().detectAndCompute returns on them: pt, size, angle,
response, packed octave field, descriptors (uint8; cv2 hands the same integers out as float32).  The smallest image
is sized for the pure-Python restatement (oracle/sift.py) to finish in seconds; the others are for the GPU path.
The script also prints how closely the restatement reproduces cv2 (the figures quoted in oracle/sift.py).

usage: python tests/golden/make_golden_sift.py      (from the repo root; needs cv2)
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import sift as S  # noqa: E402

SIZES = {"small": (96, 128, 2.0), "medium": (240, 320, 2.0), "large": (384, 512, 2.0), "odd": (93, 127, 1.6), "tiny": (30, 40, 1.2)}


def texture(h, w, sigma, seed):
    rng = np.random.default_rng(seed)
    f = cv2.GaussianBlur(rng.integers(0, 256, (h, w)).astype(np.float32), (0, 0), sigma)
    # a few large-scale blobs so that the upper octaves hold extrema as well
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for _ in range(6):
        cy, cx, s, a = rng.uniform(0, h), rng.uniform(0, w), rng.uniform(min(6, min(h, w) / 8), max(6.5, min(h, w) / 6)), rng.uniform(-25, 25)
        f += a * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s))
    return cv2.normalize(f, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)


def main():
    out = {}
    sift = cv2.SIFT_create()
    for seed, (name, (h, w, sg)) in enumerate(SIZES.items()):
        img = texture(h, w, sg, seed)
        kp, des = sift.detectAndCompute(img, None)
        arr = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response] for k in kp], np.float32)
        octv = np.array([k.octave for k in kp], np.int32)
        assert np.array_equal(des, np.rint(des)) and des.min() >= 0 and des.max() <= 255
        out[name + "_image"], out[name + "_kp"], out[name + "_octave"], out[name + "_des"] = img, arr, octv, des.astype(np.uint8)
        print("%s %dx%d: %d key points, octaves %s" % (name, w, h, len(kp), sorted(set(((o & 255) ^ 128) - 128 for o in octv))))
        if name not in ("large",):
            k2, o2, d2 = S.detect_arrays(img)
            m = S.match_keypoints(arr, k2)
            ok = m >= 0
            dd = np.abs(out[name + "_des"][ok].astype(int) - d2[m[ok]].astype(int)).max(axis=1)
            print("   restatement: %d key points, %d of cv2's reproduced (%.2f %%), octave fields equal %s, descriptors identical %d, "
                  "within 1: %d, worst %d" % (len(k2), ok.sum(), 100.0 * ok.mean(), np.array_equal(octv[ok] & 0xFFFF, o2[m[ok]] & 0xFFFF), (dd == 0).sum(),
                                              (dd <= 1).sum(), dd.max()))
    np.savez_compressed(os.path.join(HERE, "sift_reference.npz"), **out)


if __name__ == "__main__":
    main()
