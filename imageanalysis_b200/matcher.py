"""Drop-in replacement for the reference's `lib.matcher` module
(scripts/lib/matcher.py), served by the sm_100a library libiamatch.so.

Same public names, arguments and result conventions as the reference so that
`process.py` (:29, :290-292) and `3a-matching.py` keep working unchanged:

    configure()                                   matcher.py:43
    find_matches(proj, K, strategy, transform, sort, review)   :852
    raw_matches(i1, i2, k=2)                      :203
    basic_pair_matches(i1, i2)                    :218
    bidirectional_pair_matches(i1, i2, review)    :304
    ratio_pair_matches / bruteforce_pair_matches / smart_pair_matches   :595 / :696 / :358
    filter_by_transform(K, i1, i2, transform)     :90
    filter_duplicates / filter_cross_check / count_unique     :157 / :187 / :145
    saveMatches(image_list, check_if_dirty)       :1033
    globals detect_scale, the_matcher, max_distance, min_pairs, detector_node, matcher_node

Differences, all deliberate:
  * the_matcher is an exact brute-force k-NN on the GPU, not cv2's approximate
    FLANN (D1); results equal cv2.BFMatcher(norm).knnMatch bit for bit.
  * find_matches() hands the WHOLE pair work-list to one C call; the
    per-pair Python loop of the reference survives only for bookkeeping.
  * there is no CPU fallback: without libiamatch.so / a B200 the calls raise.
  * the GMS stage (cv2.xfeatures2d.matchGMS, matcher.py:285; contrib-only, D6)
    runs on the GPU (csrc/gms.cu) with the reference's arguments, so it no
    longer depends on an opencv-contrib build.
"""
from __future__ import annotations

import math
import time
from typing import List, Optional

import numpy as np

try:  # the real property tree when the reference environment is installed
    from props import getNode  # type: ignore
except ImportError:  # pragma: no cover - exercised in this repo's tests
    from .propshim import getNode

from . import _capi
from . import pairs as _pairs

from . import smart as _smart          # the pair-wise side estimators (lib.smart's API, GPU numerics)
try:
    from .logger import log, qlog  # type: ignore
except ImportError:
    def log(*args):
        print(*args)

    def qlog(*args):
        pass

detector_node = getNode('/config/detector', True)
matcher_node = getNode('/config/matcher', True)

detect_scale = 0.40
# The reference always runs the GMS grid filter inside basic_pair_matches (matcher.py:285).  False skips the stage
# (what an identity matchGMS would give); it exists for comparisons against pre-GMS fixtures, not for production.
gms_enabled = True
# Key points in the last half cell of the image: False = OpenCV's C++ matchGMS skips them in the half-cell-shifted
# grids (what cv2.xfeatures2d.matchGMS at matcher.py:285 executes); True = the wrap-around of the reference's archive
# Python restatement (scripts/lib/archive/gms_matcher.py:205), for bit-for-bit agreement with that module.
gms_archive_rule = False
# find_matches('traditional'): how many images' descriptors are brought to the host per device call, and whether images
# whose features this module loaded itself are dropped from the host once matched (the reference's cache flush,
# matcher.py:1012-1026; key points come back through Image.load_features when a later step asks for them).
host_block_images = 256
flush_host_descriptors = True
the_matcher = None
max_distance = None
min_pairs = 25

# wire `transform=` of find_matches through filter_by_transform (the reference
# accepts the argument and ignores it).  Off = as shipped.
apply_transform_filter = False
# CUDA device index used by configure(); set before calling it (one process per GPU).
device = 0

d2r = math.pi / 180.0

_norm = None


class DMatch:
    """Light stand-in for cv2.DMatch: the strategy code only reads these
    attributes (matcher.py:226-231, :487-493)."""
    __slots__ = ("queryIdx", "trainIdx", "imgIdx", "distance")

    def __init__(self, queryIdx, trainIdx, distance, imgIdx=0):
        self.queryIdx = int(queryIdx)
        self.trainIdx = int(trainIdx)
        self.imgIdx = imgIdx
        self.distance = float(distance)

    def __repr__(self):
        return "DMatch(q=%d, t=%d, d=%.3f)" % (self.queryIdx, self.trainIdx, self.distance)


class GpuKnnMatcher:
    """Object stored in `the_matcher`; exposes cv2.DescriptorMatcher.knnMatch
    (the only method the reference calls, matcher.py:212) plus the batched
    entry points find_matches() uses."""

    def __init__(self, norm: int, dev: int = 0):
        self.norm = norm
        self.device = dev
        self._engines = {}
        self._scratch = {}   # separate contexts for one-off knnMatch calls (ids 0/1 never clash with a project)

    def engine(self, desc_bytes: int) -> _capi.Engine:
        e = self._engines.get(desc_bytes)
        if e is None:
            e = _capi.Engine(self.norm, desc_bytes, self.device)
            self._engines[desc_bytes] = e
        return e

    def engine_for(self, des: np.ndarray) -> _capi.Engine:
        des = np.asarray(des)
        if des.ndim != 2:
            raise _capi.IamError("descriptors must be a 2-D array")
        return self.engine(int(des.shape[1]))

    @staticmethod
    def _as_device_dtype(des: np.ndarray, norm: int) -> np.ndarray:
        des = np.asarray(des)
        if norm == _capi.NORM_HAMMING:
            return np.ascontiguousarray(des, np.uint8)
        if des.dtype == np.uint8:
            return np.ascontiguousarray(des)
        return np.ascontiguousarray(des, np.float32)

    def knn_arrays(self, des1, des2, k):
        """(idx [N,k] int32, dist [N,k] float32) of des1 rows against des2."""
        d1 = self._as_device_dtype(des1, self.norm)
        d2 = self._as_device_dtype(des2, self.norm)
        if d1.ndim != 2 or d2.ndim != 2 or d1.shape[1] != d2.shape[1]:
            raise _capi.IamError("descriptor arrays must be [N,D] with equal D")
        eng = self._scratch.get(d1.shape[1])
        if eng is None:
            eng = self._scratch[d1.shape[1]] = _capi.Engine(self.norm, int(d1.shape[1]), self.device)
        eng.upload(0, d1)
        eng.upload(1, d2)
        n = max(d1.shape[0], 1)
        idx, dist, _, _ = eng.knn_pairs([(0, 1)], k, n, reverse=False)
        return idx[0, :d1.shape[0]], dist[0, :d1.shape[0]]

    def knnMatch(self, des1, des2, k=2):
        idx, dist = self.knn_arrays(des1, des2, k)
        out = []
        for qi in range(idx.shape[0]):
            out.append([DMatch(qi, idx[qi, s], dist[qi, s]) for s in range(k) if idx[qi, s] >= 0])
        return out


# ---------------------------------------------------------------------------
def configure():
    """matcher.py:43-80: read /config/detector and /config/matcher, pick the
    norm and max_distance, build the matcher object into `the_matcher`."""
    global detect_scale, the_matcher, max_distance, min_pairs, _norm
    detect_scale = detector_node.getFloat('scale')
    detector_str = detector_node.getString('detector')
    if detector_str == 'SIFT' or detector_str == 'SURF':
        _norm = _capi.NORM_L2
        max_distance = 270.0
    elif detector_str == 'ORB' or detector_str == 'Star':
        _norm = _capi.NORM_HAMMING
        max_distance = 64
    else:
        log("Detector not specified or not known:", detector_str)
        quit()
    the_matcher = GpuKnnMatcher(_norm, device)
    min_pairs = matcher_node.getFloat('min_pairs')


def _image_size():
    cam = getNode('/config/camera', True)
    return cam.getInt('width_px') if hasattr(cam, 'getInt') else 0, cam.getInt('height_px')


def keypoint_keys(kp_list) -> np.ndarray:
    """Integer id per keypoint such that two ids are equal iff the reference's
    string keys '%.2f-%.2f' % kp.pt (matcher.py:165-166) are equal.  kp.pt is
    float32, so pt*100 is exact in double and rint() reproduces '%.2f'."""
    if len(kp_list) == 0:
        return np.zeros((0,), np.int32)
    pts = np.asarray([kp.pt for kp in kp_list], np.float64)
    cents = np.rint(pts * 100.0).astype(np.int64)
    _, inv = np.unique(cents, axis=0, return_inverse=True)
    return inv.reshape(-1).astype(np.int32)


# ---------------------------------------------------------------------------
def filter_by_transform(K, i1, i2, transform):
    """matcher.py:90-142 with the robust fit on the GPU (batched kernel run on
    this single pair).  Removes outliers from i1.match_list[i2.name] in place
    and returns True when nothing was removed."""
    clean = True
    width = getattr(i1, 'width', None) or _image_size()[0]
    tol = math.pow(width, 0.25) if width else 1.0
    if tol < 1.0:
        tol = 1.0
    matches = i1.match_list[i2.name]
    if len(matches) < min_pairs:
        i1.match_list[i2.name] = []
        return True
    p1 = np.float32([i1.uv_list[pair[0]] for pair in matches])
    p2 = np.float32([i2.uv_list[pair[1]] for pair in matches])
    if transform == "none":
        status = np.ones(len(matches))
    elif transform in ("essential", "homography", "fundamental"):           # matcher.py:121-126
        eng = the_matcher.engine(128 if _norm == _capi.NORM_L2 else 32)
        model = {"essential": _capi.MODEL_ESSENTIAL, "homography": _capi.MODEL_HOMOGRAPHY,
                 "fundamental": _capi.MODEL_FUNDAMENTAL}[transform]
        status, M, ninl = eng.ransac_pairs(model, p1, p2, np.int32([0, len(matches)]), np.asarray(K, np.float64), tol)
    else:
        # the reference falls through to a NameError on `status` for any other string (matcher.py:129)
        raise _capi.IamError("unknown transform '%s' (homography, fundamental, essential, none)" % transform)
    log("  %s vs %s: %d / %d  inliers/matched" % (i1.name, i2.name, np.sum(status), len(status)))
    kept = []
    for k, flag in enumerate(status):
        if flag:
            kept.append(matches[k])
        else:
            clean = False
    matches[:] = kept
    return clean


def count_unique(i1, i2, matches_fit):
    """matcher.py:145-150."""
    idx_pairs = [[m.queryIdx, m.trainIdx] for m in matches_fit]
    return len(filter_duplicates(i1, i2, idx_pairs))


def filter_duplicates(i1, i2, idx_pairs):
    """matcher.py:157-182 (host version for single pairs; the batched path
    runs the same walk on the device, reduce.cu dedupe_kernel)."""
    count = 0
    result = []
    kp1_dict = {}
    kp2_dict = {}
    for pair in idx_pairs:
        kp1 = i1.kp_list[pair[0]]
        kp2 = i2.kp_list[pair[1]]
        key1 = "%.2f-%.2f" % (kp1.pt[0], kp1.pt[1])
        key2 = "%.2f-%.2f" % (kp2.pt[0], kp2.pt[1])
        if key1 in kp1_dict or key2 in kp2_dict:
            count += 1
        else:
            kp1_dict[key1] = True
            kp2_dict[key2] = True
            result.append(pair)
    if count > 0:
        qlog("  removed %d/%d duplicate features" % (count, len(idx_pairs)))
    return result


def filter_cross_check(idx_pairs1, idx_pairs2):
    """matcher.py:187-200 (set lookup instead of the O(n1*n2) list scan)."""
    have = {(int(r[0]), int(r[1])) for r in idx_pairs2}
    new1, new2 = [], []
    for pair in idx_pairs1:
        if (int(pair[1]), int(pair[0])) in have:
            new1.append(pair)
            new2.append([pair[1], pair[0]])
    if len(idx_pairs1) != len(new1) or len(idx_pairs2) != len(new2):
        qlog("  cross check: (%d, %d) => (%d, %d)" % (len(idx_pairs1), len(idx_pairs2), len(new1), len(new2)))
    return new1, new2


def raw_matches(i1, i2, k=2):
    """matcher.py:203-216."""
    if i1.des_list is None or i2.des_list is None:
        return []
    if len(i1.des_list.shape) == 0 or i1.des_list.shape[0] <= 1:
        return []
    if len(i2.des_list.shape) == 0 or i2.des_list.shape[0] <= 1:
        return []
    matches = the_matcher.knnMatch(np.array(i1.des_list), np.array(i2.des_list), k=k)
    qlog("  raw matches:", len(matches))
    return matches


def _gms(i1, i2, idx_pairs, dist_of):
    """The GMS stage of basic_pair_matches (matcher.py:275-291), on the GPU:
    matchGMS(size, size, kp1, kp2, matches, withRotation=True, withScale=False, thresholdFactor=5.0)."""
    w, h = _image_size()
    if not w or not h:
        log("Zero image sizes will crash matchGMS():", w, h)
        log("Recommend removing all meta/*.feat files and")
        log("rerun the matching step.")
        quit()
    if not idx_pairs or not gms_enabled:
        return idx_pairs
    eng = the_matcher.engine(128 if _norm == _capi.NORM_L2 else 32)
    mask = eng.gms_filter(np.float32([k.pt for k in i1.kp_list]), np.float32([k.pt for k in i2.kp_list]),
                          np.int32(idx_pairs), (w, h), with_rotation=True, with_scale=False, threshold_factor=5.0,
                          archive_wrap=gms_archive_rule)
    return [p for p, keep in zip(idx_pairs, mask) if keep]


def basic_pair_matches(i1, i2):
    """matcher.py:218-300 for one direction of one pair."""
    if i1.des_list is None or i2.des_list is None or i1.des_list.shape[0] <= 1 or i2.des_list.shape[0] <= 1:
        return []
    idx, dist = the_matcher.knn_arrays(i1.des_list, i2.des_list, 2)
    match_ratio = matcher_node.getFloat('match_ratio')
    d0 = dist[:, 0].astype(np.float64)
    d1 = dist[:, 1].astype(np.float64)
    ok = d1 != 0.0
    metric = np.full(d0.shape, np.inf)
    metric[ok] = d0[ok] * (d0[ok] / d1[ok])          # two roundings, as matcher.py:255-256
    order = np.argsort(metric, kind="stable")           # stable, as Python's sorted() (:258)
    keep = order[metric[order] < max_distance * match_ratio][:2000]   # :261, :265-269
    qlog("  quality matches:", len(keep))
    if len(keep) < min_pairs:
        return []
    idx_pairs = [[int(q), int(idx[q, 0])] for q in keep]
    idx_pairs = _gms(i1, i2, idx_pairs, lambda q: dist[q, 0])
    idx_pairs = filter_duplicates(i1, i2, idx_pairs)
    qlog("  initial matches =", len(idx_pairs))
    if len(idx_pairs) < min_pairs:
        return []
    return idx_pairs


def bidirectional_pair_matches(i1, i2, review=False):
    """matcher.py:304-347 (interactive review is not offered on this path)."""
    if i1 == i2:
        log("We shouldn't see this, but i1 == i2", i1.name, i2.name)
        return [], []
    idx_pairs1 = basic_pair_matches(i1, i2)
    if len(idx_pairs1) >= min_pairs:
        idx_pairs2 = basic_pair_matches(i2, i1)
    else:
        idx_pairs2 = []
    return filter_cross_check(idx_pairs1, idx_pairs2)


def _knn3_arrays(i1, i2, k):
    """raw_matches (matcher.py:203-216) as arrays: (idx [N,k], dist [N,k]) or None where raw_matches returns []."""
    if i1.des_list is None or i2.des_list is None:
        return None
    if len(i1.des_list.shape) == 0 or i1.des_list.shape[0] <= 1:
        return None
    if len(i2.des_list.shape) == 0 or i2.des_list.shape[0] <= 1:
        return None
    idx, dist = the_matcher.knn_arrays(np.array(i1.des_list), np.array(i2.des_list), k)
    qlog("  raw matches:", idx.shape[0])
    return idx, dist


def _kp_arrays(img):
    pts = np.float32([k.pt for k in img.kp_list]).reshape(-1, 2)
    size = np.float64([k.size for k in img.kp_list])
    angle = np.float64([k.angle for k in img.kp_list])
    return pts, size, angle


def _batched_homographies(i1, i2, groups, tol):
    """cv2.findHomography(src, dst, cv2.RANSAC, tol) (matcher.py:532, :637, :803) for every candidate group in ONE
    batched GPU call.  groups: list of (q [n], t [n]) index arrays.  Returns per group (status [n] bool, H [3,3])."""
    off = np.zeros(len(groups) + 1, np.int32)
    for n, (q, t) in enumerate(groups):
        off[n + 1] = off[n] + len(q)
    if off[-1] == 0:
        return [(np.zeros(0, bool), np.eye(3)) for _ in groups]
    pts1 = np.float32([k.pt for k in i1.kp_list]).reshape(-1, 2)
    pts2 = np.float32([k.pt for k in i2.kp_list]).reshape(-1, 2)
    p1 = np.concatenate([pts1[q] for q, _ in groups])
    p2 = np.concatenate([pts2[t] for _, t in groups])
    eng = the_matcher.engine(128 if _norm == _capi.NORM_L2 else 32)
    mask, H, _ = eng.ransac_pairs(_capi.MODEL_HOMOGRAPHY, p1, p2, off, None, float(tol))
    return [(mask[off[n]:off[n + 1]].astype(bool), H[n]) for n in range(len(groups))]


def _count_unique_idx(i1, i2, q, t):
    return len(filter_duplicates(i1, i2, [[int(a), int(b)] for a, b in zip(q, t)]))


def _finish_pairs(i1, i2, q, t, label):
    """Common tail of the strategies (matcher.py:579-593, :681-694, :836-850)."""
    if len(q) >= min_pairs:
        idx_pairs = filter_duplicates(i1, i2, [[int(a), int(b)] for a, b in zip(q, t)])
        if len(idx_pairs) >= min_pairs:
            qlog("  %s matches =" % label, len(idx_pairs))
            return idx_pairs, [[p[1], p[0]] for p in idx_pairs]
    return [], []


def _image_diag_tol(rounded):
    w, h = _image_size()
    diag = int(math.sqrt(h * h + w * w))
    tol = int(round(diag * 0.005)) if rounded else int(diag * 0.005)   # matcher.py:610 rounds, :482 / :761 truncate
    return w, h, diag, max(tol, 5)


def ratio_pair_matches(i1, i2, review=False, est_rotation=False):
    """matcher.py:595-694: cumulative bins of increasing Lowe-ratio cut-off, one homography RANSAC per bin
    (all bins in one batched GPU call), the first bin with the most unique inliers (> 20) wins."""
    knn = _knn3_arrays(i1, i2, 2)
    if knn is None:
        return [], []
    idx, dist = knn
    w, h, diag, tol = _image_diag_tol(rounded=True)
    d0, d1 = dist[:, 0].astype(np.float64), dist[:, 1].astype(np.float64)
    ok = (idx[:, 1] >= 0) & (d1 != 0.0)          # the reference divides unguarded (:620) and would raise on d1 == 0
    ratio = np.full(d0.shape, np.inf)
    ratio[ok] = d0[ok] / d1[ok]
    cutoffs = [0.5, 0.55, 0.6, 0.65, 0.7, 0.75, 0.8, 0.85]
    groups = []
    for c in cutoffs:                            # :622-630
        q = np.nonzero(ratio <= c)[0]
        groups.append((q, idx[q, 0]))
    live = [n for n, (q, _) in enumerate(groups) if len(q) >= min_pairs]
    fits = _batched_homographies(i1, i2, [groups[n] for n in live], tol)
    best_q, best_t, best_count = [], [], 20
    for n, (status, _H) in zip(live, fits):      # :632-662
        q, t = groups[n]
        fq, ft = q[status], t[status]
        uniq = _count_unique_idx(i1, i2, fq, ft)
        if uniq > best_count:
            best_q, best_t, best_count = fq, ft, uniq
    return _finish_pairs(i1, i2, best_q, best_t, "found")


def _best_of_neighbours(idx, dist, kp1, kp2, match_ratio, dist_limit, pred1=None):
    """The per-query selection shared by smart_pair_matches (:484-514) and bruteforce_pair_matches (:709-753): walk
    the k neighbours in order, stop at the first with distance >= dist_limit or d0/dj < match_ratio, skip those whose
    key-point size ratio exceeds 1.25, keep the one with the smallest metric (first on ties).
    pred1 = predicted positions of image-1 key points in image 2 (smart: metric = raw_dist * size_diff / ratio);
    None = bruteforce (metric = size_diff / ratio, raw_dist = length of the displacement p2 - p1).
    Returns (rows [m], best_j [m], raw_dist [m], vangle [m])."""
    pts1, size1, _ = kp1
    pts2, size2, _ = kp2
    n, k = idx.shape
    d = dist.astype(np.float64)
    alive = np.ones(n, bool)
    best_metric = np.full(n, np.inf)
    best_j = np.full(n, -1, np.int64)
    best_dist = np.zeros(n)
    best_vangle = np.zeros(n)
    src = pts1 if pred1 is None else pred1
    for j in range(k):
        have = idx[:, j] >= 0
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = d[:, 0] / d[:, j]
        stop = ~have | (d[:, j] >= dist_limit) | ~(ratio >= match_ratio)     # NaN (0/0) stops the walk too
        alive &= ~stop
        t = np.where(have, idx[:, j], 0)
        v = (pts2[t] - src).astype(np.float32)                               # float32 arithmetic as np.float32(kp.pt) (:725-727)
        raw = np.sqrt((v.astype(np.float32) ** 2).sum(1, dtype=np.float32)).astype(np.float64)
        s1, s2 = size1, size2[t]
        with np.errstate(divide="ignore", invalid="ignore"):
            size_diff = np.where(s1 > s2, s1 / s2, s2 / s1)
            metric = (raw * size_diff / ratio) if pred1 is not None else (size_diff / ratio)
        cand = alive & (size_diff <= 1.25)
        take = cand & ((best_j < 0) | (metric < best_metric))
        best_metric[take] = metric[take]
        best_j[take] = j
        best_dist[take] = raw[take]
        va = np.arctan2(v[:, 1].astype(np.float64), v[:, 0].astype(np.float64))
        va = np.where(va < 0, va + 2 * math.pi, va)
        best_vangle[take] = va[take]
    rows = np.nonzero(best_j >= 0)[0]
    return rows, best_j[rows], best_dist[rows], best_vangle[rows]


def bruteforce_pair_matches(i1, i2, review=False):
    """matcher.py:696-850.  k = 3 neighbours; per query the neighbour with the smallest size_diff / ratio among those
    with distance < 290, ratio >= match_ratio and size ratio <= 1.25 (:709-753); the survivors are spread into 41
    displacement-length bins x 21 displacement-direction bins, each spilling into its two neighbours (:755-795); one
    homography RANSAC per populated direction bin; the bin with the most RANSAC inliers (num_fit, from 0) wins, and the
    walk over the length bins stops early once more than 50 fitted and the current length bin gave fewer than 10
    (:797-830).  The kNN and ALL homography fits of the pair run on the GPU (one batched call); the early stop is
    replayed on the host over the batched results, so the selection is the reference's."""
    match_ratio = matcher_node.getFloat('match_ratio')
    w, h, diag, tol = _image_diag_tol(rounded=False)
    knn = _knn3_arrays(i1, i2, 3)
    if knn is None:
        return [], []
    idx, dist = knn
    kp1, kp2 = _kp_arrays(i1), _kp_arrays(i2)
    rows, bj, best_dist, best_vangle = _best_of_neighbours(idx, dist, kp1, kp2, match_ratio, 290.0)
    mq = rows
    mt = idx[rows, bj]
    maxdist = int(diag * 0.55)
    divs = 40
    step = maxdist / divs
    dist_bins = [[] for _ in range(divs + 1)]                 # :762-771 (entries = positions into mq / mt)
    for e, bd in enumerate(best_dist):
        b = int(round(bd / step))
        if b < len(dist_bins):
            dist_bins[b].append(e)
            if b > 0:
                dist_bins[b - 1].append(e)
            if b < len(dist_bins) - 1:
                dist_bins[b + 1].append(e)
    adivs = 20
    astep = 2 * math.pi / adivs
    groups, where = [], []                                    # every direction bin with >= min_pairs entries
    for bi, entries in enumerate(dist_bins):                  # :774-795
        angle_bins = [[] for _ in range(adivs + 1)]
        for e in entries:
            b = int(round(best_vangle[e] / astep))
            angle_bins[b].append(e)
            if b == 0:
                angle_bins[-1].append(e)
                angle_bins[b + 1].append(e)
            elif b == adivs:
                angle_bins[b - 1].append(e)
                angle_bins[0].append(e)
            else:
                angle_bins[b - 1].append(e)
                angle_bins[b + 1].append(e)
        for ab in angle_bins:
            if len(ab) >= min_pairs:
                ee = np.asarray(ab, np.int64)
                groups.append((mq[ee], mt[ee]))
                where.append(bi)
    fits = _batched_homographies(i1, i2, groups, tol) if groups else []
    best_fitted, fit_q, fit_t = 0, [], []
    g = 0
    for bi in range(len(dist_bins)):                          # :797-830, replayed in order over the batched fits
        best_of_bin = 0
        while g < len(groups) and where[g] == bi:
            status, _H = fits[g]
            num_fit = int(np.count_nonzero(status))
            best_of_bin = max(best_of_bin, num_fit)
            if num_fit > best_fitted:
                fit_q, fit_t = groups[g][0][status], groups[g][1][status]
                best_fitted = num_fit
            g += 1
        if best_fitted > 50 and best_of_bin < 10:
            break
    return _finish_pairs(i1, i2, fit_q, fit_t, "initial")


def _fit_homography_lsq(src, dst):
    """cv2.findHomography(src, dst, 0) (matcher.py:452): all-point least squares -- normalised DLT (Hartley)."""
    src = np.asarray(src, np.float64).reshape(-1, 2)
    dst = np.asarray(dst, np.float64).reshape(-1, 2)

    def norm(p):
        c = p.mean(0)
        s = math.sqrt(2.0) / max(np.sqrt(((p - c) ** 2).sum(1)).mean(), 1e-12)
        T = np.array([[s, 0, -s * c[0]], [0, s, -s * c[1]], [0, 0, 1.0]])
        return (p - c) * s, T
    a, Ta = norm(src)
    b, Tb = norm(dst)
    A = np.zeros((2 * len(a), 9))
    A[0::2, 0:2], A[0::2, 2] = a, 1.0
    A[0::2, 6:8], A[0::2, 8] = -b[:, :1] * a, -b[:, 0]
    A[1::2, 3:5], A[1::2, 5] = a, 1.0
    A[1::2, 6:8], A[1::2, 8] = -b[:, 1:] * a, -b[:, 1]
    _, _, vt = np.linalg.svd(A)
    H = np.linalg.inv(Tb) @ vt[-1].reshape(3, 3) @ Ta
    return H / H[2, 2] if abs(H[2, 2]) > 1e-12 else H


def _project_points(pts_ned, R, tvec, K, dist):
    """cv2.projectPoints (matcher.py:435): pinhole + (k1, k2, p1, p2, k3) distortion."""
    X = (np.asarray(pts_ned, np.float64) @ np.asarray(R, np.float64).T) + np.asarray(tvec, np.float64).reshape(1, 3)
    x, y = X[:, 0] / X[:, 2], X[:, 1] / X[:, 2]
    k1, k2, p1, p2, k3 = (list(np.asarray(dist, np.float64).ravel()) + [0.0] * 5)[:5]
    r2 = x * x + y * y
    rad = 1 + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2
    xd = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    return np.stack([K[0, 0] * xd + K[0, 2], K[1, 1] * yd + K[1, 2]], 1)


# smart_pair_matches: predicted homography image 1 -> image 2.  A caller may install its own predictor
# (i1, i2, est_rotation) -> 3x3 here; None = the reference's pose-based prediction when the images carry poses
# (get_body2ned / get_cam2body / get_camera_pose), else the identity.
predict_homography = None


def _quat_matrix(q):
    """transformations.quaternion_matrix (used by image.py:537-539), rotation part."""
    q = np.asarray(q, np.float64)
    n = np.dot(q, q)
    if n < np.finfo(float).eps * 4.0:
        return np.identity(3)
    q = q * math.sqrt(2.0 / n)
    q = np.outer(q, q)
    return np.array([[1.0 - q[2, 2] - q[3, 3], q[1, 2] - q[3, 0], q[1, 3] + q[2, 0]],
                     [q[1, 2] + q[3, 0], 1.0 - q[1, 1] - q[3, 3], q[2, 3] - q[1, 0]],
                     [q[1, 3] - q[2, 0], q[2, 3] + q[1, 0], 1.0 - q[1, 1] - q[2, 2]]])


def _rot_x(angle_rad):
    c, s = math.cos(angle_rad), math.sin(angle_rad)
    return np.array([[1.0, 0, 0], [0, c, -s], [0, s, c]])      # rotation_matrix(angle, [1, 0, 0])[:3, :3]


def _pose_prediction(i1, i2, est_rotation):
    """matcher.py:359-454: project an 8 x 8-step grid of image-2 pixels onto the assumed ground plane, re-project the
    3-D points into image 1 and fit the homography image 1 -> image 2 through the 81 correspondences."""
    need = ("get_body2ned", "get_cam2body", "get_camera_pose")
    if not all(hasattr(i1, a) and hasattr(i2, a) for a in need):
        return None
    cam = getNode('/config/camera', True)
    K = np.array([cam.getFloatEnum('K', i) for i in range(9)], np.float64).reshape(3, 3)     # camera.get_K()
    if abs(K[0, 0]) < 1e-9:
        return None
    dist = [cam.getFloatEnum('dist_coeffs', i) for i in range(5)] if cam.hasChild('dist_coeffs') else [0.0] * 5
    w, h = _image_size()
    IK = np.linalg.inv(K)
    steps = 8
    grid = np.array([[u, v] for v in np.linspace(0, h, steps + 1) for u in np.linspace(0, w, steps + 1)])   # gen_grid :349-356
    if matcher_node.hasChild("ground_m"):
        ground_m = matcher_node.getFloat("ground_m")
    elif _smart is not None:
        ground_m = _smart.get_surface_estimate(i1, i2)
    else:
        return None
    y1 = _smart.get_yaw_error_estimate(i1) if _smart is not None else 0.0
    y2 = _smart.get_yaw_error_estimate(i2) if _smart is not None else 0.0
    if abs(y1) < 0.0001 and abs(y2) > 0.0001:      # :385-388
        y1 = y2
    if abs(y1) > 0.0001 and abs(y2) < 0.0001:
        y2 = y1
    body2ned2 = np.asarray(i2.get_body2ned(), np.float64)
    if est_rotation:
        body2ned2 = body2ned2 @ _rot_x(y2 * d2r)
    cam2body2 = np.asarray(i2.get_cam2body(), np.float64)
    uvh = np.concatenate([grid, np.ones((len(grid), 1))], 1)
    rays = (body2ned2 @ cam2body2 @ IK @ uvh.T).T                                            # project.projectVectors :536-548
    rays /= np.linalg.norm(rays, axis=1, keepdims=True)
    ned2 = np.asarray(i2.get_camera_pose()[0], np.float64)
    if -ned2[2] < ground_m:
        ground_m = -ned2[2] - 2
    pts = np.tile(ned2, (len(rays), 1))                                                     # intersectVectorsWithGroundPlane :553-565
    down = rays[:, 2] > 0.0
    factor = -(ned2[2] + ground_m) / rays[down, 2]
    pts[down] = ned2 + rays[down] * factor[:, None]
    body2ned1 = np.asarray(i1.get_body2ned(), np.float64)                                   # Image.get_proj, image.py:543-553
    if est_rotation and abs(y1) > 0.001:
        body2ned1 = body2ned1 @ _rot_x(y1 * d2r)
    R = np.asarray(i1.get_cam2body(), np.float64).T @ body2ned1.T
    ned1 = np.asarray(i1.get_camera_pose()[0], np.float64)
    reproj = _project_points(pts, R, -R @ ned1, K, dist)
    return _fit_homography_lsq(reproj, grid)


def smart_pair_matches(i1, i2, review=False, est_rotation=False):
    """matcher.py:358-593.  Start from a homography predicted from the two camera poses and the ground estimate
    (:359-454), then iterate (:472-577): move image-1 key points through the current homography, pick per query the
    neighbour (k = 3) with the smallest  predicted-position error x size ratio / Lowe ratio, bin the survivors by that
    error (cumulative cut-offs 32 ... 2048 px), fit one homography per bin with RANSAC and keep the bin with the most
    unique inliers; repeat from its homography until no bin improves.  The kNN and every round's homography fits
    (one batched call per round) run on the GPU; the prediction comes from `predict_homography` when installed, else
    from the poses the images carry (same arithmetic as lib.project / Image.get_proj), else it is the identity."""
    match_ratio = matcher_node.getFloat("match_ratio")
    w, h, diag, tol = _image_diag_tol(rounded=False)
    H = None
    if predict_homography is not None:
        H = predict_homography(i1, i2, est_rotation)
    if H is None:
        H = _pose_prediction(i1, i2, est_rotation)
    if H is None:
        log("  smart: no pose information on", i1.name, i2.name, "-- starting from the identity")
        H = np.eye(3)
    knn = _knn3_arrays(i1, i2, 3)
    if knn is None:
        return [], []
    idx, dist = knn
    kp1, kp2 = _kp_arrays(i1), _kp_arrays(i2)
    src = kp1[0].astype(np.float64)
    best_count, best_q, best_t = 20, [], []
    cutoffs = [32, 64, 128, 256, 512, 1024, 2048]
    while True:
        ph = np.concatenate([src, np.ones((len(src), 1))], 1) @ np.asarray(H, np.float64).T     # cv2.perspectiveTransform :474
        with np.errstate(divide="ignore", invalid="ignore"):
            trans = (ph[:, :2] / ph[:, 2:3]).astype(np.float32)
        rows, bj, best_dist, _ = _best_of_neighbours(idx, dist, kp1, kp2, match_ratio, 300.0, pred1=trans)
        mq, mt = rows, idx[rows, bj]
        groups = []
        for c in cutoffs:                                   # :519-524
            e = np.nonzero(best_dist < c)[0]
            groups.append((mq[e], mt[e]))
        live = [n for n, (q, _) in enumerate(groups) if len(q) >= min_pairs]
        fits = _batched_homographies(i1, i2, [groups[n] for n in live], tol)
        done = True
        for n, (status, H_test) in zip(live, fits):         # :526-560
            q, t = groups[n]
            fq, ft = q[status], t[status]
            uniq = _count_unique_idx(i1, i2, fq, ft)
            if uniq > best_count:
                done = False
                H = np.array(H_test, np.float64)
                best_q, best_t, best_count = fq, ft, uniq
        if done:
            break
    return _finish_pairs(i1, i2, best_q, best_t, "found")


# ---------------------------------------------------------------------------
def _ensure_features(img):
    if img.kp_list is None or img.des_list is None or not len(img.kp_list) or not len(img.des_list):
        img.detect_features(detect_scale)


def find_matches(proj, K, strategy="smart", transform="homography", sort=False, review=False):
    """matcher.py:852-1031.  Work-list generation, resume rules, result
    storage and saving follow the reference; the per-pair matching of the
    'traditional' strategy runs as ONE batched GPU call over the whole list."""
    image_list = proj.image_list
    neds = [im.get_camera_pose()[0] for im in image_list]
    min_dist = matcher_node.getFloat("min_dist") if matcher_node.hasChild("min_dist") else 0
    max_dist = matcher_node.getFloat("max_dist") if matcher_node.hasChild("max_dist") else None
    mode = matcher_node.getString("pair_filter") if matcher_node.hasChild("pair_filter") else "sequential"
    work_list = _pairs.worklist(neds, mode=mode or "sequential", min_dist=min_dist, max_dist=max_dist)
    log('Generating work list for range:', min_dist, '-', max_dist)
    if sort:
        work_list = sorted(work_list, key=lambda fields: fields[0])   # matcher.py:916

    todo = []
    for dist, i, j in work_list:
        i1, i2 = image_list[i], image_list[j]
        if i2.name in i1.match_list and i1.name in i2.match_list:     # matcher.py:946-951
            if len(i1.match_list[i2.name]) == 0:
                log("Retrying: ", i1.name, "vs", i2.name, "(no matches found previously)")
            else:
                log("Skipping: ", i1.name, "vs", i2.name, "already done.")
                continue
        todo.append((i, j))
    log("Processing worklist matches:")

    if strategy == "traditional":
        results = _batched_traditional(image_list, todo)
    else:
        results = []
        for i, j in todo:
            i1, i2 = image_list[i], image_list[j]
            _ensure_features(i1)
            _ensure_features(i2)
            if strategy == "smart":
                results.append(smart_pair_matches(i1, i2, review, True))
            elif strategy == "bestratio":
                results.append(ratio_pair_matches(i1, i2, review, True))
            elif strategy == "bruteforce":
                results.append(bruteforce_pair_matches(i1, i2))
            else:
                raise ValueError("unknown strategy '%s'" % strategy)

    for (i, j), (match_fwd, match_rev) in zip(todo, results):
        i1, i2 = image_list[i], image_list[j]
        i1.desc_timestamp = time.time()
        i2.desc_timestamp = time.time()
        i1.match_list[i2.name] = match_fwd                           # matcher.py:979-984
        i2.match_list[i1.name] = match_rev
        i1.matches_clean = False
        i2.matches_clean = False
        if apply_transform_filter and len(match_fwd):
            filter_by_transform(K, i1, i2, transform)
            i2.match_list[i1.name] = [[p[1], p[0]] for p in i1.match_list[i2.name]]
        if _smart is not None and all(hasattr(im, a) for im in (i1, i2) for a in ("get_proj", "get_aircraft_pose", "set_aircraft_yaw_error_estimate")):
            avg, std = _smart.update_surface_estimate(i1, i2)         # matcher.py:987-1005
            i1.set_aircraft_yaw_error_estimate(_smart.update_yaw_error_estimate(i1, i2))
            i2.set_aircraft_yaw_error_estimate(_smart.update_yaw_error_estimate(i2, i1))
            if std and std >= 50 and len(i1.match_list[i2.name]) < 100:
                log("Std dev of surface triangulation blew up, matches are probably bad so discarding them!",
                    i1.name, i2.name, "avg:", avg, "std:", std, "count:", len(match_fwd))
                i1.match_list[i2.name] = []
                i2.match_list[i1.name] = []

    saveMatches(image_list)
    if _smart is not None:
        _smart.save(proj.analysis_dir)
    print('Pair-wise matches successfully saved.')


def _batched_traditional(image_list, todo):
    """bidirectional_pair_matches for every pair of `todo` in one device
    pipeline: kNN both ways -> metric reduction -> [GMS] -> filter_duplicates
    -> min_pairs gates -> cross-check (reduce.cu).

    The pair list is walked in blocks that bring at most `host_block_images` NEW images to the host at a time: their
    descriptors are loaded (Image.detect_features / the .desc cache), handed to one iam_match_images call together with
    the block's pairs (images of earlier blocks stay resident on the device), and -- when this function loaded them --
    dropped from the host again right after the upload (later pairs read the device copy), which is what the reference's descriptor-cache flush
    (matcher.py:1012-1026) is for: the float32 descriptors of a 2812-frame project are 7.2 GB, the device copies 1.8 GB."""
    if not todo:
        return []
    eng = None
    w, h = _image_size()
    if not w or not h:
        log("Zero image sizes will crash matchGMS():", w, h)
        quit()
    prm = _capi.Engine.make_params(match_ratio=matcher_node.getFloat('match_ratio'), max_distance=float(max_distance),
                                   reduce_mode=_capi.REDUCE_REF_METRIC, cap=2000, min_pairs=int(min_pairs),
                                   cross_check=True, dedupe=True, gms=(2 if gms_archive_rule else 1) if gms_enabled else 0,
                                   gms_rotation=True, gms_scale=False,
                                   gms_threshold=5.0, size=(w, h))
    # The call uploads an image right before the first pair that needs it: handing it the pairs sorted by their LATER
    # image lets matching start after the first few images instead of after everything the first image's neighbours
    # reach (results are independent of the order, matcher.py:928-980; they are mapped back to the work-list order).
    order = sorted(range(len(todo)), key=lambda p: (max(todo[p]), min(todo[p])))
    out = [None] * len(todo)
    uploaded, ours = set(), set()
    pos = 0
    while pos < len(order):
        new, end = [], pos
        while end < len(order):
            need = [i for i in sorted(set(todo[order[end]])) if i not in uploaded and i not in new]
            if new and len(new) + len(need) > max(2, int(host_block_images)):
                break
            new += need
            end += 1
        for i in new:
            img = image_list[i]
            if img.des_list is None or img.kp_list is None or not len(img.kp_list):
                ours.add(i)
            _ensure_features(img)
        arrays = [GpuKnnMatcher._as_device_dtype(image_list[i].des_list, _norm) for i in new]
        if len({a.dtype for a in arrays}) > 1:   # mixed caches: widen everything to float32
            arrays = [np.ascontiguousarray(a, np.float32) for a in arrays]
        keys = [keypoint_keys(image_list[i].kp_list) for i in new]
        if eng is None:
            eng = the_matcher.engine(int(arrays[0].shape[1]))
        if gms_enabled:   # keypoint coordinates for the GMS stage (matcher.py:285); the descriptors follow inside the call
            eng.upload_keypoints_batch(new, [np.float32([k.pt for k in image_list[i].kp_list]).reshape(-1, 2) for i in new])
        table, count = eng.match_images(new, arrays, np.int32([todo[p] for p in order[pos:end]]).reshape(-1, 2), prm, keys=keys)
        for k, p in enumerate(order[pos:end]):
            fwd = table[k, :count[k]].tolist()
            out[p] = (fwd, [[t, q] for q, t in fwd])
        uploaded.update(new)
        del arrays, keys, table, count
        if flush_host_descriptors and not apply_transform_filter:
            for i in [i for i in new if i in ours]:   # later pairs read the device copy: the host copy is not needed again
                img = image_list[i]
                img.kp_list = None
                img.des_list = None
                img.uv_list = None
                ours.discard(i)
        pos = end
    return out


def saveMatches(image_list, check_if_dirty=False):
    """matcher.py:1033-1040."""
    log('saving matches and image meta data ...')
    for image in image_list:
        if check_if_dirty:
            if not image.matches_clean:
                image.save_matches()
        else:
            image.save_matches()


# -- small helpers the reference exports (visualisation-free parts) ---------
def decomposeAffine(affine):
    """matcher.py:1043-1065."""
    tx = affine[0][2]
    ty = affine[1][2]
    a, b = affine[0][0], affine[0][1]
    c, d = affine[1][0], affine[1][1]
    sx = math.sqrt(a * a + b * b)
    if a < 0.0:
        sx = -sx
    sy = math.sqrt(c * c + d * d)
    if d < 0.0:
        sy = -sy
    rotate_deg = math.atan2(-b, a) * 180.0 / math.pi
    if rotate_deg < -180.0:
        rotate_deg += 360.0
    if rotate_deg > 180.0:
        rotate_deg -= 360.0
    return (rotate_deg, tx, ty, sx, sy)


def copyKeyPoint(k):
    """matcher.py:1095-1099."""
    import cv2
    return cv2.KeyPoint(x=k.pt[0], y=k.pt[1], size=k.size, angle=k.angle, response=k.response, octave=k.octave,
                        class_id=k.class_id)


class Matcher():
    """Legacy container instantiated at import by lib/match_cleanup.py:17 and
    3b-clean-and-combine-matches.py:19; its review/plot methods are GUI code
    outside the accelerated path."""

    def __init__(self):
        self.image_list = []

    def setImageList(self, image_list):
        self.image_list = image_list

    def findImageIndex(self, search):
        for i, image in enumerate(self.image_list):
            if search == image:
                return i
        return None

    def findImageByName(self, search):
        for i, image in enumerate(self.image_list):
            if search == image.name:
                return image
        return None


def group_matches(matches_direct):
    raise NotImplementedError("group_matches is broken in the reference snapshot (matcher.py:1736) and unused")
