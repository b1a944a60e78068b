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

try:  # sibling reference modules when this file is dropped into scripts/lib/
    from . import smart as _smart  # type: ignore
except ImportError:
    _smart = None
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
    elif transform in ("essential", "homography"):
        eng = the_matcher.engine(128 if _norm == _capi.NORM_L2 else 32)
        model = _capi.MODEL_ESSENTIAL if transform == "essential" else _capi.MODEL_HOMOGRAPHY
        status, M, ninl = eng.ransac_pairs(model, p1, p2, np.int32([0, len(matches)]), np.asarray(K, np.float64), tol)
    else:
        raise _capi.IamError("transform '%s' has no GPU implementation (essential, homography, none)" % transform)
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
                          np.int32(idx_pairs), (w, h), with_rotation=True, with_scale=False, threshold_factor=5.0)
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


def _homography_bins(i1, i2, bins, tol):
    """Shared tail of ratio_pair_matches / bruteforce_pair_matches: fit a
    homography per candidate bin on the GPU (all bins in ONE batched call)
    and keep the bin with the most unique inliers (matcher.py:632-662, :797-830)."""
    best, best_count = [], 20
    live = [b for b in bins if len(b) >= min_pairs]
    if not live:
        return []
    off = np.zeros(len(live) + 1, np.int32)
    p1, p2 = [], []
    for n, b in enumerate(live):
        off[n + 1] = off[n] + len(b)
        p1.extend(i1.kp_list[m.queryIdx].pt for m in b)
        p2.extend(i2.kp_list[m.trainIdx].pt for m in b)
    eng = the_matcher.engine(128 if _norm == _capi.NORM_L2 else 32)
    mask, H, ninl = eng.ransac_pairs(_capi.MODEL_HOMOGRAPHY, np.float32(p1), np.float32(p2), off, None, float(tol))
    for n, b in enumerate(live):
        fit = [m for m, f in zip(b, mask[off[n]:off[n + 1]]) if f]
        uniq = count_unique(i1, i2, fit)
        if uniq > best_count:
            best, best_count = fit, uniq
    return best


def _finish_bins(i1, i2, matches_best):
    if len(matches_best) >= min_pairs:
        idx_pairs = filter_duplicates(i1, i2, [[m.queryIdx, m.trainIdx] for m in matches_best])
        if len(idx_pairs) >= min_pairs:
            qlog("  found matches =", len(idx_pairs))
            return idx_pairs, [[p[1], p[0]] for p in idx_pairs]
    return [], []


def ratio_pair_matches(i1, i2, review=False, est_rotation=False):
    """matcher.py:595-694: bins of increasing Lowe-ratio cut-off, one
    homography RANSAC per bin, best bin by unique inliers."""
    matches = raw_matches(i1, i2, k=2)
    w, h = _image_size()
    diag = int(math.sqrt(h * h + w * w))
    tol = max(5, int(round(diag * 0.005)))
    cutoffs = [0.5, 0.55, 0.6, 0.65, 0.7, 0.75, 0.8, 0.85]
    bins = [[] for _ in cutoffs]
    for m in matches:
        if len(m) < 2 or m[1].distance == 0:
            continue
        ratio = m[0].distance / m[1].distance
        for i, c in enumerate(cutoffs):
            if ratio <= c:
                bins[i].append(m[0])
    return _finish_bins(i1, i2, _homography_bins(i1, i2, bins, tol))


def bruteforce_pair_matches(i1, i2, review=False):
    """matcher.py:696-850: k=3 neighbours, binned by displacement length and
    direction, one homography RANSAC per populated bin."""
    match_ratio = matcher_node.getFloat('match_ratio')
    w, h = _image_size()
    diag = int(math.sqrt(h * h + w * w))
    tol = max(5, int(round(diag * 0.005)))
    matches = raw_matches(i1, i2, k=3)
    dist_steps, angle_steps = 8, 8
    bins = [[[] for _ in range(angle_steps)] for _ in range(dist_steps)]
    for m in matches:
        for j in range(len(m)):
            if j + 1 < len(m) and not (m[j].distance <= m[j + 1].distance * match_ratio) and j > 0:
                continue
            p1 = i1.kp_list[m[j].queryIdx].pt
            p2 = i2.kp_list[m[j].trainIdx].pt
            dx, dy = p2[0] - p1[0], p2[1] - p1[1]
            di = min(dist_steps - 1, int(math.hypot(dx, dy) / max(diag, 1) * dist_steps))
            ai = int(((math.atan2(dy, dx) + math.pi) / (2 * math.pi)) * angle_steps) % angle_steps
            bins[di][ai].append(m[j])
    flat = [b for row in bins for b in row]
    return _finish_bins(i1, i2, _homography_bins(i1, i2, flat, tol))


def smart_pair_matches(i1, i2, review=False, est_rotation=False):
    """matcher.py:358-593 needs the reference's pose/SRTM machinery
    (lib.smart, lib.srtm, camera mounts) to predict feature locations; that
    is outside the accelerated path.  The k=3 neighbour
    search it starts from (:465) is `raw_matches(i1, i2, k=3)`."""
    raise NotImplementedError(
        "strategy 'smart' depends on lib.smart/lib.srtm pose prediction which this module does not "
        "re-implement; use strategy 'traditional', 'bestratio' or 'bruteforce'")


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
        if _smart is not None:                                        # matcher.py:987-1005
            avg, std = _smart.update_surface_estimate(i1, i2)
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
    -> min_pairs gates -> cross-check (reduce.cu)."""
    if not todo:
        return []
    used = sorted({i for p in todo for i in p})
    for i in used:
        _ensure_features(image_list[i])
    arrays = [GpuKnnMatcher._as_device_dtype(image_list[i].des_list, _norm) for i in used]
    if len({a.dtype for a in arrays}) > 1:   # mixed caches: widen everything to float32
        arrays = [np.ascontiguousarray(a, np.float32) for a in arrays]
    keys = [keypoint_keys(image_list[i].kp_list) for i in used]
    eng = the_matcher.engine(int(arrays[0].shape[1]))
    w, h = _image_size()
    if not w or not h:
        log("Zero image sizes will crash matchGMS():", w, h)
        quit()
    if gms_enabled:   # keypoint coordinates for the GMS stage (matcher.py:285); the descriptors follow inside the call
        for i in used:
            eng.upload_keypoints(i, np.float32([k.pt for k in image_list[i].kp_list]).reshape(-1, 2))
    prm = _capi.Engine.make_params(match_ratio=matcher_node.getFloat('match_ratio'), max_distance=float(max_distance),
                                   reduce_mode=_capi.REDUCE_REF_METRIC, cap=2000, min_pairs=int(min_pairs),
                                   cross_check=True, dedupe=True, gms=gms_enabled, gms_rotation=True, gms_scale=False,
                                   gms_threshold=5.0, size=(w, h))
    # one C call: uploads are enqueued wave by wave so PCIe overlaps the matching
    table, count = eng.match_images(used, arrays, np.int32(todo), prm, keys=keys)
    out = []
    for p in range(len(todo)):
        fwd = table[p, :count[p]].tolist()
        out.append((fwd, [[t, q] for q, t in fwd]))
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
