"""Drop-in for the feature detector the reference builds in Image.make_detector
(scripts/lib/image.py:230-251) and calls in Image.detect_features (:324):

    detector = cv2.SIFT_create()                  # the default, :236-237
    detector = cv2.ORB_create(max_features)       # :243-245
    self.kp_list, self.des_list = detector.detectAndCompute(scaled, None)

`SIFT_create()` returns an object with the same call that builds the Gaussian / difference-of-Gaussian pyramid, finds
and refines the scale-space extrema, assigns orientations and computes the 128-byte descriptors on the GPU
(csrc/sift.cu through iam_sift_detect).

`ORB_create(n)` here returns an object with the same `detectAndCompute(image, mask)` call that runs FAST-9, the
Harris ranking, the intensity-centroid orientation and the steered BRIEF descriptors of all 8 pyramid levels on the
GPU (csrc/orb.cu through iam_orb_detect) and hands back key points with cv2.KeyPoint's attributes (real
cv2.KeyPoint objects when cv2 is importable, so that the reference's cache writer, image.py:192-207, is served
unchanged) and the uint8 [N, 32] descriptor array.  There is no CPU fallback.
"""
from __future__ import annotations

import queue

import numpy as np

from . import _capi

_engine = None
_pool = []
device = 0


def _eng():
    global _engine
    if _engine is None:
        _engine = _capi.Engine(_capi.NORM_HAMMING, 32, device)
    return _engine


def to_gray(image: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(image, COLOR_BGR2GRAY) as ORB applies it to colour input: fixed-point
    (B*1868 + G*9617 + R*4899 + 8192) >> 14."""
    image = np.asarray(image)
    if image.ndim == 2:
        return np.ascontiguousarray(image, np.uint8)
    if image.ndim == 3 and image.shape[2] in (3, 4):
        b, g, r = (image[:, :, i].astype(np.int32) for i in range(3))
        return ((b * 1868 + g * 9617 + r * 4899 + 8192) >> 14).astype(np.uint8)
    raise _capi.IamError("image must be [H, W] or [H, W, 3|4] uint8")


def sift_detect_and_compute(image: np.ndarray) -> dict:
    """Arrays: pt [n, 2] f32, size, angle, response [n] f32, octave [n] i32 (cv2's packed field), des [n, 128] u8."""
    kp, octv, des = _eng().sift_detect(to_gray(image))
    return dict(pt=kp[:, 0:2].copy(), size=kp[:, 2].copy(), angle=kp[:, 3].copy(), response=kp[:, 4].copy(), octave=octv, des=des)


def sift_detect_many(images, workers: int = 2):
    """sift_detect_and_compute for a list of images with `workers` contexts in flight (one host thread, stream and
    pyramid block each): while one frame's key points are sorted on the host and its descriptors cross PCIe, the next
    frame's pyramid is built.  Results are in the order of `images` and identical to the one-by-one calls."""
    from concurrent.futures import ThreadPoolExecutor
    images = list(images)
    if workers <= 1 or len(images) <= 1:
        return [sift_detect_and_compute(im) for im in images]
    while len(_pool) < workers:                       # contexts (and their pyramid blocks) are kept between calls
        _pool.append(_capi.Engine(_capi.NORM_HAMMING, 32, device))
    free = queue.Queue()
    for e in _pool[:workers]:
        free.put(e)

    def one(im):
        eng = free.get()
        try:
            kp, octv, des = eng.sift_detect(to_gray(im))
        finally:
            free.put(eng)
        return dict(pt=kp[:, 0:2].copy(), size=kp[:, 2].copy(), angle=kp[:, 3].copy(), response=kp[:, 4].copy(), octave=octv, des=des)

    with ThreadPoolExecutor(max_workers=workers) as pool:
        return list(pool.map(one, images))


def orb_detect_and_compute(image: np.ndarray, nfeatures: int = 500) -> dict:
    """Arrays: pt [n, 2] f32, size, angle, response [n] f32, octave [n] i32, des [n, 32] u8."""
    kp, des = _eng().orb_detect(to_gray(image), nfeatures)
    return dict(pt=kp[:, 0:2].copy(), size=kp[:, 2].copy(), angle=kp[:, 3].copy(), response=kp[:, 4].copy(),
                octave=kp[:, 5].astype(np.int32), des=des)


class _KeyPoint:
    __slots__ = ("pt", "size", "angle", "response", "octave", "class_id")

    def __init__(self, x, y, size, angle, response, octave):
        self.pt = (float(x), float(y))
        self.size = float(size)
        self.angle = float(angle)
        self.response = float(response)
        self.octave = int(octave)
        self.class_id = -1


def _keypoints(r):
    try:
        import cv2
        mk = lambda x, y, s, a, rs, o: cv2.KeyPoint(x=float(x), y=float(y), size=float(s), angle=float(a),  # noqa: E731
                                                    response=float(rs), octave=int(o), class_id=-1)
    except ImportError:
        mk = _KeyPoint
    return [mk(p[0], p[1], s, a, rs, o) for p, s, a, rs, o in zip(r["pt"], r["size"], r["angle"], r["response"], r["octave"])]


class ORB:
    def __init__(self, nfeatures: int = 500):
        self.nfeatures = int(nfeatures)

    def detectAndCompute(self, image, mask=None):
        if mask is not None:
            raise _capi.IamError("masks are not supported (the reference passes None, image.py:324)")
        r = orb_detect_and_compute(image, self.nfeatures)
        kps = _keypoints(r)
        return kps, (r["des"] if len(kps) else None)


class SIFT:
    """cv2.SIFT_create() with OpenCV's defaults; `uint8_descriptors=True` hands the descriptors out as the uint8 the
    kernel produces (cv2 returns the same integers as float32; the matcher narrows them back, matcher.py:123)."""

    def __init__(self, uint8_descriptors: bool = False):
        self.uint8_descriptors = bool(uint8_descriptors)

    def detectAndCompute(self, image, mask=None):
        if mask is not None:
            raise _capi.IamError("masks are not supported (the reference passes None, image.py:324)")
        r = sift_detect_and_compute(image)
        kps = _keypoints(r)
        if not kps:
            return kps, None
        return kps, (r["des"] if self.uint8_descriptors else r["des"].astype(np.float32))


def SIFT_create(uint8_descriptors: bool = False) -> SIFT:
    """cv2.SIFT_create() (image.py:236-237): the reference passes no arguments."""
    return SIFT(uint8_descriptors)


def ORB_create(nfeatures: int = 500) -> ORB:
    """cv2.ORB_create(nfeatures) (image.py:244-245) with OpenCV's default parameters."""
    return ORB(nfeatures)


def make_detector(name: str = "SIFT", max_features: int = 20000):
    """The dispatch of Image.make_detector (image.py:230-251) on /config/detector/detector: 'SIFT' (built without
    arguments there) and 'ORB' (orb_max_features) run on the GPU; 'SURF' and 'Star' are cv2.xfeatures2d detectors that
    this framework does not replace."""
    if name == "SIFT":
        return SIFT_create()
    if name == "ORB":
        return ORB_create(max_features)
    raise _capi.IamError("detector %r is not replaced by this framework (SIFT and ORB are)" % (name,))
