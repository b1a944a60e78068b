"""The on-disk feature cache of the reference, written and read in its formats (scripts/lib/image.py):

  <project>/ImageAnalysis/cache/<image>.feat   gzip(pickle([((x, y), size, angle, response, octave, class_id), ...]))
                                                 -- Image.save_features :192-207, load_features :139-158
  <project>/ImageAnalysis/cache/<image>.desc   gzip(np.save(des_list)): float32 [N, 128] (SIFT) or uint8 [N, 32] (ORB)
                                                 -- Image.save_descriptors :209-217, load_descriptors :160-180

`detect_and_cache` is the tail of Image.detect_features (:324, :343-349) on top of the GPU detectors: detect on the scaled
image, move the key points back to full-resolution pixels, write both files.  Files written here load with the reference's
own Image.load_features / load_descriptors and vice versa (tests/test_featcache.py).
"""
from __future__ import annotations

import gzip
import pickle

import numpy as np

from . import detector as _detector


def save_features(path: str, kp_list) -> None:
    feature_list = [((float(kp.pt[0]), float(kp.pt[1])), kp.size, kp.angle, kp.response, kp.octave, kp.class_id) for kp in kp_list]
    with gzip.open(path, "wb", compresslevel=6) as fp:
        pickle.dump(feature_list, fp)


def load_features(path: str):
    """Key points as cv2.KeyPoint objects when cv2 is importable (what Image.load_features builds), else objects with
    the same attributes."""
    with gzip.open(path, "rb") as fp:
        feature_list = pickle.load(fp)
    rows = dict(pt=[p[0] for p in feature_list], size=[p[1] for p in feature_list], angle=[p[2] for p in feature_list],
                response=[p[3] for p in feature_list], octave=[p[4] for p in feature_list])
    kps = _detector._keypoints(rows)
    for kp, p in zip(kps, feature_list):
        kp.class_id = p[5]
    return kps


def save_descriptors(path: str, des_list: np.ndarray) -> None:
    with gzip.open(path, "wb", compresslevel=6) as fp:
        np.save(fp, des_list)


def load_descriptors(path: str) -> np.ndarray:
    with gzip.open(path, "rb") as fp:
        return np.load(fp)


def detect_and_cache(scaled_image: np.ndarray, scale: float, feat_path: str, desc_path: str, det=None):
    """detector.detectAndCompute(scaled, None) (:324), kp.pt /= scale (:343-346), save_features + save_descriptors
    (:348-349).  `det` defaults to the GPU SIFT (the reference's default detector).  Returns (kp_list, des_list)."""
    det = det or _detector.SIFT_create()
    kp_list, des_list = det.detectAndCompute(scaled_image, None)
    for kp in kp_list:
        kp.pt = (kp.pt[0] / scale, kp.pt[1] / scale)
    save_features(feat_path, kp_list)
    save_descriptors(desc_path, des_list if des_list is not None else np.zeros((0, 128), np.float32))
    return kp_list, des_list
