#!/usr/bin/env python3
"""Golden fixtures for the ORB detect + describe stage (reference scripts/lib/image.py:243-245, :324:
`cv2.ORB_create(max_features).detectAndCompute(scaled, None)`), recorded from LIVE cv2 in the build container.

1. orb_pattern.npy -- the 256 binary tests (x0, y0, x1, y1) of OpenCV's learned BRIEF pattern.  OpenCV's source is
   not available offline, so the table is RECOVERED from the library itself by probing `cv2.ORB.compute` with one
   key point (angle 0, octave 0) at the centre of step images: for a vertical edge at column c the descriptor blur
   (7 taps) makes the profile strictly increasing over [c-3, c+3] and flat outside, so bit i = (I(p0) < I(p1)) is set
   exactly for c in [x0-2, x1+3] when x0 < x1 (falling edges give the pairs with x0 > x1; horizontal edges the y
   coordinates; pairs with equal x (or y) are resolved with a corner image that lights only the half plane holding
   the second point).  The table is then verified on random images through the descriptor restatement.
2. orb_reference.npz -- two test images and what cv2.ORB_create(n).detectAndCompute returns on them
   (n = 500 and 2000): pt, size, angle, response, octave, descriptors.

usage: python tests/golden/make_golden_orb.py      (from the repo root; needs cv2)
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
S, R = 121, 20
C = S // 2


def _desc(orb, img):
    kp = [cv2.KeyPoint(x=float(C), y=float(C), size=31.0, angle=0.0, response=1.0, octave=0, class_id=-1)]
    _, d = orb.compute(img, kp)
    return np.unpackbits(d[0], bitorder="little")     # bit k of byte j = test 8 j + k


def recover_pattern():
    orb = cv2.ORB_create(500)
    P = np.full((256, 4), 99, np.int64)
    for axis in (0, 1):
        runs = {}
        for rising in (True, False):
            rows = []
            for c in range(-R, R + 1):
                img = np.zeros((S, S), np.uint8)
                sl = slice(C + c, None) if rising else slice(None, C + c)
                if axis == 0:
                    img[:, sl] = 255
                else:
                    img[sl, :] = 255
                rows.append(_desc(orb, img))
            runs[rising] = np.array(rows)
        for i in range(256):
            cu, cd = np.nonzero(runs[True][:, i])[0], np.nonzero(runs[False][:, i])[0]
            assert not (len(cu) and len(cd))
            if len(cu):      # first < second along this axis
                P[i, axis], P[i, 2 + axis] = cu[0] - R + 2, cu[-1] - R - 3
            elif len(cd):
                P[i, 2 + axis], P[i, axis] = cd[0] - R + 2, cd[-1] - R - 3
    for i in range(256):     # equal coordinates along one axis: corner images
        for axis in (0, 1):
            if P[i, axis] != 99:
                continue
            o = 1 - axis
            v0, v1 = P[i, o], P[i, 2 + o]
            assert v0 != 99 and v0 != v1
            last = None
            for c in range(-R, R + 1):
                lit_o = np.zeros(S, bool)
                if v0 < v1:
                    lit_o[C + v0 + 1:] = True
                else:
                    lit_o[:C + v1 + 1] = True
                lit_a = np.zeros(S, bool)
                lit_a[C + c:] = True
                img = np.zeros((S, S), np.uint8)
                if axis == 0:
                    img[np.ix_(lit_o, lit_a)] = 255
                else:
                    img[np.ix_(lit_a, lit_o)] = 255
                if _desc(orb, img)[i]:
                    last = c
            P[i, axis] = P[i, 2 + axis] = last - 3
    assert (np.abs(P) <= 15).all()
    # verification: the restated descriptor (float separable blur, angle 0) on random images
    k = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    rng = np.random.default_rng(0)
    bad = 0
    for _ in range(200):
        img = cv2.GaussianBlur(rng.integers(0, 256, (S, S)).astype(np.uint8), (0, 0), 1.5)
        b = cv2.sepFilter2D(img, cv2.CV_8U, k, k, borderType=cv2.BORDER_REFLECT_101)
        got = (b[C + P[:, 1], C + P[:, 0]] < b[C + P[:, 3], C + P[:, 2]]).astype(np.uint8)
        bad += int((got != _desc(orb, img)).sum())
    print("pattern recovered; mismatching bits on 200 random images:", bad)
    assert bad == 0
    np.save(os.path.join(HERE, "orb_pattern.npy"), P.astype(np.int8))


def texture(seed, w=640, h=480, sigma=2.0):
    rng = np.random.default_rng(seed)
    img = cv2.GaussianBlur(rng.integers(0, 256, (h, w)).astype(np.float32), (0, 0), sigma)
    return cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)


def blocks(seed, w=600, h=440):
    """Rectangles and discs on a gradient: strong corners, flat regions, FAST score ties."""
    rng = np.random.default_rng(seed)
    img = np.tile(np.linspace(40, 200, w, dtype=np.float32), (h, 1))
    for _ in range(60):
        x, y = int(rng.integers(0, w - 40)), int(rng.integers(0, h - 40))
        a, b = int(rng.integers(8, 60)), int(rng.integers(8, 60))
        v = int(rng.integers(0, 256))
        if rng.random() < 0.5:
            cv2.rectangle(img, (x, y), (x + a, y + b), v, -1)
        else:
            cv2.circle(img, (x + 20, y + 20), a // 2 + 3, v, -1)
    img = cv2.GaussianBlur(img, (0, 0), 0.8) + rng.normal(0, 1.5, (h, w)).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    recover_pattern()
    out = {}
    for name, img in (("texture", texture(5)), ("blocks", blocks(6))):
        out[name + "_img"] = img
        for n in (500, 2000):
            kps, des = cv2.ORB_create(n).detectAndCompute(img, None)
            tag = "%s_%d" % (name, n)
            out[tag + "_pt"] = np.float32([k.pt for k in kps])
            out[tag + "_size"] = np.float32([k.size for k in kps])
            out[tag + "_angle"] = np.float32([k.angle for k in kps])
            out[tag + "_response"] = np.float32([k.response for k in kps])
            out[tag + "_octave"] = np.int32([k.octave for k in kps])
            out[tag + "_des"] = des
            print(tag, len(kps), "key points")
        fk = cv2.FastFeatureDetector_create(20, True).detect(img)
        out[name + "_fast"] = np.int32(sorted((int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in fk))
    np.savez_compressed(os.path.join(HERE, "orb_reference.npz"), **out)


if __name__ == "__main__":
    main()
