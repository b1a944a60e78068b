#!/usr/bin/env python3
"""Golden fixtures for the GMS grid filter (matcher.py:285).

cv2.xfeatures2d (contrib) is absent from this image, so the pin is the
reference's own restatement of that algorithm, scripts/lib/archive/gms_matcher.py,
imported UNMODIFIED from /root/reference and run with the threshold factor the
reference's call site passes (thresholdFactor=5.0; the module constant is 6).

Key points of the main scenes are kept inside the first 97 % of the image: for a
point in the last half cell the archive module indexes mCellPairs[-1] (Python
wrap-around to cell 399) where the C++ original skips the match.  The CUDA kernel
follows the C++ behaviour (it replaces the cv2 call); the oracle restates both
(archive_wrap=True is the literal Python) and the "edge_strip" scene pins the
archive behaviour on exactly that strip.

usage: python tests/golden/make_golden_gms.py      (from the repo root; needs /root/reference)
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/scripts/lib/archive")
import gms_matcher as ref  # noqa: E402

ref.THRESHOLD_FACTOR = 5.0          # matcher.py:285 thresholdFactor=5.0


def scene(n, inlier_frac, angle_deg, scale, seed, w=5472, h=3648, clustered=False):
    rng = np.random.default_rng(seed)
    if clustered:   # matches concentrated in a few cells: large per-cell counts
        centres = rng.uniform(0.2, 0.75, (6, 2))
        p1 = centres[rng.integers(0, 6, n)] + rng.normal(0, 0.02, (n, 2))
        p1 = np.clip(p1, 0.01, 0.96)
    else:
        p1 = rng.uniform(0.0, 0.97, (n, 2))
    a = np.deg2rad(angle_deg)
    R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
    p2 = (p1 - 0.5) @ R.T * scale + 0.5 + rng.normal(0, 0.002, (n, 2))
    out = rng.random(n) > inlier_frac
    p2[out] = rng.uniform(0.0, 0.97, (int(out.sum()), 2))
    ok = (p2 >= 0).all(1) & (p2 < 0.97).all(1)
    p1, p2 = p1[ok], p2[ok]
    n = len(p1)
    # key-point tables are larger than the match list and shuffled, as in the reference (matches index into them)
    n1, n2 = n + 37, n + 11
    pts1 = rng.uniform(0, 0.97, (n1, 2))
    pts2 = rng.uniform(0, 0.97, (n2, 2))
    qi = rng.permutation(n1)[:n]
    ti = rng.permutation(n2)[:n]
    pts1[qi] = p1
    pts2[ti] = p2
    pts1 = (pts1 * [w, h]).astype(np.float32)     # cv2.KeyPoint.pt is float32
    pts2 = (pts2 * [w, h]).astype(np.float32)
    return pts1, pts2, np.stack([qi, ti], 1).astype(np.int32), (w, h)


def run_ref(pts1, pts2, matches, size, with_scale, with_rotation):
    dm = [types.SimpleNamespace(queryIdx=int(q), trainIdx=int(t)) for q, t in matches]
    s = ref.Size(size[0], size[1])
    with contextlib.redirect_stdout(io.StringIO()):
        g = ref.GmsMatcher([tuple(map(float, p)) for p in pts1], s, [tuple(map(float, p)) for p in pts2], s, dm)
        mask, n_in = g.GetInlierMask(with_scale, with_rotation)
    return np.array(mask, bool), int(n_in)


def gen_pipeline():
    """The reference's own lib.matcher (unmodified) with cv2.xfeatures2d.matchGMS served by the reference's
    archive GmsMatcher: basic_pair_matches / bidirectional_pair_matches and the find_matches driver with the GMS
    stage live (matcher.py:285), on a strip of synthetic frames whose planted matches move by a common shift."""
    import cv2
    REPO = os.path.dirname(os.path.dirname(HERE))
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(HERE, "shims"))
    sys.path.insert(0, "/root/reference/scripts")
    import make_golden as mg
    from imageanalysis_b200 import synth
    matcher, _ = mg.import_reference_matcher()

    def match_gms(size1, size2, kp1, kp2, matches, withRotation=False, withScale=False, thresholdFactor=6.0):
        ref.THRESHOLD_FACTOR = thresholdFactor
        with contextlib.redirect_stdout(io.StringIO()):
            g = ref.GmsMatcher([k.pt for k in kp1], ref.Size(*size1), [k.pt for k in kp2], ref.Size(*size2), matches)
            mask, _ = g.GetInlierMask(withScale, withRotation)
        return [m for m, keep in zip(matches, mask) if keep]

    cv2.xfeatures2d = types.SimpleNamespace(matchGMS=match_gms)
    n = 5
    des, pts, neds = synth.sift_project(n, 1200, seed=17, planted=0.4)
    for p in pts:                       # stay clear of the last half cell (module docstring)
        p[:, 0] = np.minimum(p[:, 0], 0.97 * 5472)
        p[:, 1] = np.minimum(p[:, 1], 0.97 * 3648)
        p[40:60] = p[0:20]              # colliding keypoints: filter_duplicates has work after GMS
    imgs = [mg.FakeImage("frame%02d" % i, des[i].astype(np.float32), pts[i], neds[i]) for i in range(n)]
    out = {"n": n}
    fwd = matcher.basic_pair_matches(imgs[0], imgs[1])
    rev = matcher.basic_pair_matches(imgs[1], imgs[0])
    out["basic01"] = np.int32(fwd).reshape(-1, 2)
    out["basic10"] = np.int32(rev).reshape(-1, 2)
    b1, b2 = matcher.bidirectional_pair_matches(imgs[1], imgs[2])
    out["bidir12_fwd"] = np.int32(b1).reshape(-1, 2)
    out["bidir12_rev"] = np.int32(b2).reshape(-1, 2)
    print("pipeline: basic01", len(fwd), "basic10", len(rev), "bidir12", len(b1))
    proj = types.SimpleNamespace(image_list=imgs, analysis_dir="/tmp")
    K = np.array([[3666.666504, 0, 2736], [0, 3666.666504, 1824], [0, 0, 1]])
    with contextlib.redirect_stdout(io.StringIO()):
        matcher.find_matches(proj, K, strategy="traditional", transform="homography", sort=False, review=False)
    for i, im in enumerate(imgs):
        out["des%d" % i] = des[i]
        out["pts%d" % i] = pts[i]
        out["ned%d" % i] = np.float64(neds[i])
        for other, lst in im.match_list.items():
            out["match_%s_%s" % (im.name, other)] = np.int32(lst).reshape(-1, 2)
            print("  ", im.name, other, len(lst))
    np.savez_compressed(os.path.join(HERE, "reference_gms_pipeline.npz"), **out)


def main():
    cases = {
        "rot0": dict(n=1200, inlier_frac=0.6, angle_deg=0, scale=1.0, seed=1),
        "rot90": dict(n=2000, inlier_frac=0.5, angle_deg=90, scale=0.9, seed=2),
        "rot180": dict(n=800, inlier_frac=0.7, angle_deg=180, scale=1.0, seed=3),
        "rot45": dict(n=1500, inlier_frac=0.6, angle_deg=45, scale=0.8, seed=4),
        "rot225_small": dict(n=120, inlier_frac=0.8, angle_deg=225, scale=1.0, seed=5),
        "outliers_only": dict(n=400, inlier_frac=0.0, angle_deg=0, scale=1.0, seed=6),
        "clustered": dict(n=2000, inlier_frac=0.8, angle_deg=10, scale=1.0, seed=7, clustered=True),
    }
    out = {}
    for name, kw in cases.items():
        pts1, pts2, matches, size = scene(**kw)
        mask, n_in = run_ref(pts1, pts2, matches, size, False, True)     # the reference's flags (matcher.py:285)
        print(name, len(matches), "matches ->", n_in, "inliers")
        out[name + "_pts1"], out[name + "_pts2"], out[name + "_matches"] = pts1, pts2, matches
        out[name + "_size"] = np.array(size, np.int32)
        out[name + "_mask"] = mask
    # the other flag combinations on one scene
    pts1, pts2, matches, size = scene(n=900, inlier_frac=0.6, angle_deg=0, scale=0.55, seed=8)
    for ws, wr in ((False, False), (True, False), (True, True)):
        mask, n_in = run_ref(pts1, pts2, matches, size, ws, wr)
        print("flags", ws, wr, "->", n_in)
        out["flags_s%d_r%d_mask" % (ws, wr)] = mask
    out["flags_pts1"], out["flags_pts2"], out["flags_matches"], out["flags_size"] = pts1, pts2, matches, np.array(size, np.int32)
    # Key points in the LAST HALF CELL of the image (no cell in the shifted grids): here the archive module wraps
    # around to cell 399 (mCellPairs[-1], gms_matcher.py:205) where OpenCV's C++ skips the match.  The scene puts a
    # cluster of consistent matches into cell 399 -> the same right cell, so that the wrap-around really admits
    # edge-strip matches the C++ rule drops: the fixture pins the archive behaviour, tests/test_gms.py documents both.
    rng = np.random.default_rng(9)
    n = 900
    p1 = rng.uniform(0.05, 0.9, (n, 2))
    p2 = p1 + rng.normal(0, 0.002, (n, 2))
    p1[:120] = rng.uniform(0.955, 0.972, (120, 2))              # cell 399 (x, y in [0.95, 1)) minus its last half
    p2[:120] = p1[:120] - 0.3 + rng.normal(0, 0.001, (120, 2))  # all land in one right cell
    p1[120:160, 0] = rng.uniform(0.976, 0.999, 40)              # x in the last half cell: index -1 in grids 2 and 4
    p1[120:160, 1] = rng.uniform(0.3, 0.6, 40)
    p2[120:160] = p2[:40]                                       # ... paired with the right cell that cell 399 maps to
    w, h = 2000, 1000
    pts1 = (p1 * [w, h]).astype(np.float32)
    pts2 = (p2 * [w, h]).astype(np.float32)
    matches = np.stack([np.arange(n), np.arange(n)], 1).astype(np.int32)
    mask, n_in = run_ref(pts1, pts2, matches, (w, h), False, True)
    print("edge_strip", n, "matches ->", n_in, "inliers;", int(mask[120:160].sum()), "of the 40 edge-strip matches kept by the wrap-around")
    out["edge_strip_pts1"], out["edge_strip_pts2"], out["edge_strip_matches"] = pts1, pts2, matches
    out["edge_strip_size"] = np.array((w, h), np.int32)
    out["edge_strip_mask"] = mask
    out["names"] = np.array(sorted(cases))
    np.savez_compressed(os.path.join(HERE, "gms_reference.npz"), **out)
    gen_pipeline()


if __name__ == "__main__":
    main()
