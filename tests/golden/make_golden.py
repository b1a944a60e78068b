#!/usr/bin/env python3
"""Generate the golden fixtures that pin the oracle (and through it the CUDA
path) to the reference's real behaviour.

Runs ONLY in the build container (needs /root/reference and cv2); the GPU box
just reads the committed .npz files.  Two sources of truth are recorded:

  1. live OpenCV: cv2.BFMatcher(norm).knnMatch — the exact matcher the
     reference's call site `the_matcher.knnMatch(...)` (scripts/lib/matcher.py:212)
     approximates with FLANN and uses directly elsewhere
     (scripts/lib/find_obj.py:46);
  2. the reference's own Python, imported unmodified from
     /root/reference/scripts/lib/matcher.py with the shims in ./shims:
     basic_pair_matches (:218-300), bidirectional_pair_matches (:304-347),
     filter_cross_check (:187-200), filter_duplicates (:157-182) and the whole
     find_matches driver (:852-1031) on a tiny project.  cv2.xfeatures2d.matchGMS
     (contrib-only, absent here: SURVEY D6) is replaced by an identity pass
     that also records its input, which is exactly the metric-sorted,
     thresholded, clipped list of matcher.py:253-269.

usage: python tests/golden/make_golden.py      (from the repo root)
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, os.path.join(REF, "scripts", "lib", "archive"))  # transformations.py (Gohlke), used in place
sys.path.insert(0, os.path.join(REF, "scripts"))

import cv2  # noqa: E402

from imageanalysis_b200 import synth  # noqa: E402


def bf_knn(q, t, norm, k):
    bf = cv2.BFMatcher(norm)
    m = bf.knnMatch(q, t, k=k)
    idx = np.full((len(m), k), -1, np.int32)
    dist = np.full((len(m), k), np.inf, np.float32)
    for i, row in enumerate(m):
        for s, dm in enumerate(row):
            idx[i, s] = dm.trainIdx
            dist[i, s] = dm.distance
    return idx, dist


def gen_knn_l2_synth():
    rng = np.random.default_rng(1)
    t = synth.sift_like(300, seed=11)
    q = synth.sift_like(384, seed=12)
    q[:120] = np.clip(t[rng.permutation(300)[:120]].astype(np.int32) + rng.integers(-3, 4, (120, 128)), 0, 255)
    # adversarial rows: exact duplicates (ties), zeros, saturation
    t[200:220] = t[100:120]          # duplicated train rows -> equal distances, lower index must win
    t[250] = 0
    t[251] = 255
    q[300] = 0
    q[301] = 255
    q[302] = t[100]                  # exact hit with a duplicate further down
    qf, tf = q.astype(np.float32), t.astype(np.float32)
    idx, dist = bf_knn(qf, tf, cv2.NORM_L2, 3)
    ridx, rdist = bf_knn(tf, qf, cv2.NORM_L2, 3)
    np.savez_compressed(os.path.join(HERE, "knn_l2_synth.npz"), q=q, t=t, idx=idx, dist=dist, ridx=ridx, rdist=rdist)


def gen_knn_hamming_synth():
    rng = np.random.default_rng(2)
    t = synth.orb_like(260, seed=21)
    q = synth.orb_like(300, seed=22)
    q[:100] = synth.flip_bits(t[rng.permutation(260)[:100]], 20, rng)
    t[200:210] = t[50:60]
    t[255] = 0
    t[256] = 255
    q[290] = 0
    q[291] = 255
    q[292] = t[50]
    idx, dist = bf_knn(q, t, cv2.NORM_HAMMING, 3)
    ridx, rdist = bf_knn(t, q, cv2.NORM_HAMMING, 3)
    np.savez_compressed(os.path.join(HERE, "knn_hamming_synth.npz"), q=q, t=t, idx=idx, dist=dist, ridx=ridx,
                        rdist=rdist)


def texture_pair(seed=0, w=800, h=600):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w)).astype(np.float32)
    img = cv2.GaussianBlur(img, (0, 0), 2.0)
    img = cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
    M = cv2.getRotationMatrix2D((w / 2, h / 2), 7.0, 1.0)
    M[0, 2] += 60
    M[1, 2] -= 25
    img2 = cv2.warpAffine(img, M, (w, h), borderMode=cv2.BORDER_REFLECT)
    return img, img2


def gen_real():
    a, b = texture_pair()
    sift = cv2.SIFT_create(nfeatures=900)
    kp1, d1 = sift.detectAndCompute(a, None)
    kp2, d2 = sift.detectAndCompute(b, None)
    assert (d1 == np.rint(d1)).all() and d1.max() <= 255 and d1.min() >= 0
    i12, s12 = bf_knn(d1, d2, cv2.NORM_L2, 3)
    i21, s21 = bf_knn(d2, d1, cv2.NORM_L2, 3)
    np.savez_compressed(os.path.join(HERE, "knn_sift_real.npz"), q=d1.astype(np.uint8), t=d2.astype(np.uint8),
                        idx=i12, dist=s12, ridx=i21, rdist=s21,
                        pts1=np.float32([k.pt for k in kp1]), pts2=np.float32([k.pt for k in kp2]))
    orb = cv2.ORB_create(700)
    ko1, o1 = orb.detectAndCompute(a, None)
    ko2, o2 = orb.detectAndCompute(b, None)
    i12, s12 = bf_knn(o1, o2, cv2.NORM_HAMMING, 3)
    i21, s21 = bf_knn(o2, o1, cv2.NORM_HAMMING, 3)
    np.savez_compressed(os.path.join(HERE, "knn_orb_real.npz"), q=o1, t=o2, idx=i12, dist=s12, ridx=i21, rdist=s21,
                        pts1=np.float32([k.pt for k in ko1]), pts2=np.float32([k.pt for k in ko2]))


# ---------------------------------------------------------------------------
# the reference module itself
# ---------------------------------------------------------------------------
class FakeImage:
    """Duck-typed stand-in for lib.image.Image (image.py:25-97): only the
    attributes lib.matcher touches."""

    def __init__(self, name, des, pts, ned):
        self.name = name
        self.des_list = des
        self.kp_list = [cv2.KeyPoint(x=float(p[0]), y=float(p[1]), size=4.0) for p in pts]
        self.uv_list = [list(map(float, p)) for p in pts]
        self.match_list = {}
        self.matches_clean = True
        self.desc_timestamp = 0.0
        self._ned = list(map(float, ned))

    def get_camera_pose(self, opt=False):
        return self._ned, [0.0, 0.0, 0.0], [1.0, 0.0, 0.0, 0.0]

    def detect_features(self, scale):
        raise AssertionError("descriptors are preloaded")

    def save_matches(self):
        self.matches_clean = True

    def set_aircraft_yaw_error_estimate(self, v):
        pass


def import_reference_matcher():
    from props import getNode
    det = getNode("/config/detector", True)
    det.setString("detector", "SIFT")
    det.setFloat("scale", 1.0)
    mn = getNode("/config/matcher", True)
    mn.setFloat("match_ratio", 0.75)
    mn.setFloat("min_pairs", 25)
    cam = getNode("/config/camera", True)
    cam.setInt("width_px", 5472)
    cam.setInt("height_px", 3648)
    from lib import matcher, smart  # the unmodified reference module
    captured = {}

    def gms_identity(size1, size2, kp1, kp2, matches, **kw):
        captured["thresh"] = [[m.queryIdx, m.trainIdx] for m in matches]
        return matches

    cv2.xfeatures2d = types.SimpleNamespace(matchGMS=gms_identity)
    smart.update_surface_estimate = lambda i1, i2: (None, None)
    smart.update_yaw_error_estimate = lambda i1, i2: 0.0
    smart.save = lambda d: None
    matcher.configure()
    matcher.the_matcher = cv2.BFMatcher(cv2.NORM_L2)  # exact NN in place of FLANN
    matcher.camera.get_image_params = lambda: (5472, 3648)
    return matcher, captured


def gen_reference_reductions(matcher, captured):
    des, pts, neds = synth.sift_project(4, 1500, seed=5, planted=0.4)
    # make some keypoints collide at 2-decimal precision so filter_duplicates has work to do
    for p in pts:
        p[40:60] = p[0:20]
    imgs = [FakeImage("img%03d" % i, des[i].astype(np.float32), pts[i], neds[i]) for i in range(4)]
    out = {}
    fwd = matcher.basic_pair_matches(imgs[0], imgs[1])
    out["thresh01"] = np.int32(captured["thresh"])
    out["basic01"] = np.int32(fwd).reshape(-1, 2)
    rev = matcher.basic_pair_matches(imgs[1], imgs[0])
    out["thresh10"] = np.int32(captured["thresh"])
    out["basic10"] = np.int32(rev).reshape(-1, 2)
    c1, c2 = matcher.filter_cross_check(fwd, rev)
    out["cross01"] = np.int32(c1).reshape(-1, 2)
    out["cross10"] = np.int32(c2).reshape(-1, 2)
    b1, b2 = matcher.bidirectional_pair_matches(imgs[0], imgs[2])
    out["bidir02_fwd"] = np.int32(b1).reshape(-1, 2)
    out["bidir02_rev"] = np.int32(b2).reshape(-1, 2)
    # a pair with no overlap: must come back empty (min_pairs gate, matcher.py:271-273)
    far = FakeImage("far", synth.sift_like(1500, seed=999).astype(np.float32), pts[0], neds[0])
    e1, e2 = matcher.bidirectional_pair_matches(imgs[0], far)
    out["bidir_far_fwd"] = np.int32(e1).reshape(-1, 2)
    out["bidir_far_rev"] = np.int32(e2).reshape(-1, 2)
    for i in range(4):
        out["des%d" % i] = des[i]
        out["pts%d" % i] = pts[i]
    out["des_far"] = far.des_list.astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "reference_reductions.npz"), **out)


def gen_reference_find_matches(matcher):
    """The whole driver (matcher.py:852-1031) on a 7-image strip, 'traditional'
    strategy (process.py:80-81 default), sequential |i-j|<=4 work list."""
    n = 7
    des, pts, neds = synth.sift_project(n, 1200, seed=9, planted=0.4)
    imgs = [FakeImage("frame%02d" % i, des[i].astype(np.float32), pts[i], neds[i]) for i in range(n)]
    proj = types.SimpleNamespace(image_list=imgs, analysis_dir="/tmp")
    K = np.array([[3666.666504, 0, 2736], [0, 3666.666504, 1824], [0, 0, 1]])
    matcher.find_matches(proj, K, strategy="traditional", transform="homography", sort=False, review=False)
    out = {"n": n}
    for i, im in enumerate(imgs):
        out["des%d" % i] = des[i]
        out["pts%d" % i] = pts[i]
        out["ned%d" % i] = np.float64(neds[i])
        for other, lst in im.match_list.items():
            out["match_%s_%s" % (im.name, other)] = np.int32(lst).reshape(-1, 2)
    np.savez_compressed(os.path.join(HERE, "reference_find_matches.npz"), **out)


def gen_findessential():
    """cv2.findEssentialMat exactly as matcher.py:126 calls it, on synthetic two-view scenes."""
    K = np.array([[3666.666504, 0, 2736], [0, 3666.666504, 1824], [0, 0, 1]])
    tol = max(1.0, 5472 ** 0.25)
    out = {"K": K, "tol": tol}
    for s, (n, frac) in enumerate([(2000, 0.3), (400, 0.5), (60, 0.1)]):
        p1, p2, truth = synth.two_view_scene(n, frac, K, seed=100 + s)
        E, mask = cv2.findEssentialMat(p1, p2, K, cv2.RANSAC, threshold=tol)
        out["p1_%d" % s] = p1
        out["p2_%d" % s] = p2
        out["truth_%d" % s] = truth
        out["E_%d" % s] = E[:3]
        out["mask_%d" % s] = mask.ravel().astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "find_essential.npz"), **out)


if __name__ == "__main__":
    gen_knn_l2_synth()
    gen_knn_hamming_synth()
    gen_real()
    m, cap = import_reference_matcher()
    gen_reference_reductions(m, cap)
    gen_reference_find_matches(m)
    gen_findessential()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
