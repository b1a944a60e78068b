#!/usr/bin/env python3
"""BASELINE configs[0]: "16 synthetic 1920x1080 frames, SIFT, BFMatcher L2 + ratio via the reference's matching step
on CPU (plumbing, no GPU)".

Sixteen overlapping 1920 x 1080 frames are cut out of one smooth random texture (a strip flown left to right, ~78 %
overlap between neighbours, a little yaw), features are detected as Image.detect_features does it
(scripts/lib/image.py:287-350: `cv2.resize(rgb, (0,0), fx=scale, fy=scale)` with the default --scale 0.4,
`cv2.SIFT_create().detectAndCompute(scaled, None)`, key points scaled back to full resolution), and the UNMODIFIED
reference `lib.matcher.find_matches` -- the function scripts/3a-matching.py:112 and process.py:291 call -- runs the
'traditional' strategy over its sequential work list with the exact cv2.BFMatcher injected and
cv2.xfeatures2d.matchGMS served by the reference's own archive GmsMatcher.  Recorded: every image's
`match_list` (what Image.save_matches pickles into meta/<name>.match, image.py:219-228) plus a digest of the
features.  The frames and features are NOT stored (60 MB): tests/test_config0.py regenerates them with the same code
(`frames()` / `features()` below, needs cv2 on the test box) and checks the digest before comparing.

usage: python tests/golden/make_golden_config0.py      (from the repo root; needs /root/reference and cv2)
"""
import contextlib
import hashlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
W, H, N, SCALE = 1920, 1080, 16, 0.4


def frames():
    import cv2
    rng = np.random.default_rng(16)
    world = cv2.GaussianBlur(rng.integers(0, 256, (2400, 9200)).astype(np.float32), (0, 0), 7.0)
    world = cv2.normalize(world, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
    out = []
    for f in range(N):
        cx, cy, ang = 1150 + f * 420, 1200 + 10 * np.sin(f), 2.0 * np.sin(0.7 * f)
        M = cv2.getRotationMatrix2D((cx, cy), ang, 1.0)
        M[0, 2] += W / 2 - cx
        M[1, 2] += H / 2 - cy
        out.append(cv2.warpAffine(world, M, (W, H), flags=cv2.INTER_LINEAR))
    return out


def features(frame):
    """Image.detect_features (image.py:306-346) for the SIFT detector."""
    import cv2
    scaled = cv2.resize(frame, (0, 0), fx=SCALE, fy=SCALE)
    kp, des = cv2.SIFT_create().detectAndCompute(scaled, None)
    pts = np.float32([(k.pt[0] / SCALE, k.pt[1] / SCALE) for k in kp])
    return pts, des


def digest(feats):
    h = hashlib.sha256()
    for pts, des in feats:
        h.update(np.ascontiguousarray(pts).tobytes())
        h.update(np.ascontiguousarray(des).tobytes())
    return h.hexdigest()


def main():
    import cv2
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    sys.path.insert(0, os.path.join(HERE, "shims"))
    sys.path.insert(0, "/root/reference/scripts/lib/archive")
    sys.path.insert(0, "/root/reference/scripts")
    import gms_matcher as gms
    import make_golden as mg
    matcher, _ = mg.import_reference_matcher()
    matcher.camera.get_image_params = lambda: (W, H)

    def match_gms(size1, size2, kp1, kp2, matches, withRotation=False, withScale=False, thresholdFactor=6.0):
        gms.THRESHOLD_FACTOR = thresholdFactor
        with contextlib.redirect_stdout(io.StringIO()):
            g = gms.GmsMatcher([k.pt for k in kp1], gms.Size(*size1), [k.pt for k in kp2], gms.Size(*size2), matches)
            mask, _ = g.GetInlierMask(withScale, withRotation)
        return [m for m, keep in zip(matches, mask) if keep]

    cv2.xfeatures2d = types.SimpleNamespace(matchGMS=match_gms)
    feats = [features(f) for f in frames()]
    print("features per frame:", [len(p) for p, _ in feats])
    imgs = [mg.FakeImage("frame%02d" % i, des, pts, (0.0, 12.0 * i, -60.0)) for i, (pts, des) in enumerate(feats)]
    proj = types.SimpleNamespace(image_list=imgs, analysis_dir="/tmp")
    K = np.array([[1388.0, 0, 960.0], [0, 1388.0, 540.0], [0, 0, 1]])
    with contextlib.redirect_stdout(io.StringIO()):
        matcher.find_matches(proj, K, strategy="traditional", transform="homography", sort=False, review=False)
    out = {"digest": digest(feats), "n": N, "counts": np.int32([len(p) for p, _ in feats])}
    total = 0
    for im in imgs:
        for other, lst in im.match_list.items():
            out["match_%s_%s" % (im.name, other)] = np.int32(lst).reshape(-1, 2)
            total += len(lst)
    print("pairs stored:", sum(len(im.match_list) for im in imgs), "matches:", total)
    np.savez_compressed(os.path.join(HERE, "reference_config0.npz"), **out)


if __name__ == "__main__":
    main()
