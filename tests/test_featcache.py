"""The .feat / .desc cache files (reference scripts/lib/image.py:139-217): written here, read the way the reference's
Image.load_features / load_descriptors read them, and the other way round."""
import gzip
import pickle

import numpy as np
import pytest

from imageanalysis_b200 import detector, featcache


def _kps(n, seed=0):
    rng = np.random.default_rng(seed)
    rows = dict(pt=rng.uniform(0, 4000, (n, 2)).astype(np.float32), size=rng.uniform(2, 30, n).astype(np.float32),
                angle=rng.uniform(0, 360, n).astype(np.float32), response=rng.uniform(0, 0.1, n).astype(np.float32),
                octave=rng.integers(0, 1 << 23, n).astype(np.int32))
    return detector._keypoints(rows)


def test_feat_file_is_what_the_reference_reads(tmp_path):
    kps = _kps(50)
    path = str(tmp_path / "img.feat")
    featcache.save_features(path, kps)
    fp = gzip.open(path, "rb")                      # Image.load_features, image.py:143-151
    feature_list = pickle.load(fp)
    fp.close()
    assert len(feature_list) == 50
    for point, kp in zip(feature_list, kps):
        assert len(point) == 6 and isinstance(point[0], tuple)
        assert point[0][0] == kp.pt[0] and point[0][1] == kp.pt[1] and point[1] == kp.size and point[2] == kp.angle
        assert point[3] == kp.response and point[4] == kp.octave and point[5] == kp.class_id == -1
    back = featcache.load_features(path)
    assert [(k.pt, k.size, k.angle, k.response, k.octave, k.class_id) for k in back] == \
           [(k.pt, k.size, k.angle, k.response, k.octave, k.class_id) for k in kps]


def test_feat_file_written_the_reference_way_loads(tmp_path):
    kps = _kps(20, seed=1)
    path = str(tmp_path / "ref.feat")
    feature_list = [(kp.pt, kp.size, kp.angle, kp.response, kp.octave, kp.class_id) for kp in kps]   # Image.save_features :194-199
    fp = gzip.open(path, "wb", compresslevel=6)
    pickle.dump(feature_list, fp)
    fp.close()
    back = featcache.load_features(path)
    assert [k.pt for k in back] == [k.pt for k in kps] and [k.octave for k in back] == [k.octave for k in kps]


@pytest.mark.parametrize("dtype,width", [(np.float32, 128), (np.uint8, 32)])
def test_desc_file_round_trip(tmp_path, dtype, width):
    des = np.random.default_rng(2).integers(0, 200, (77, width)).astype(dtype)
    path = str(tmp_path / "img.desc")
    featcache.save_descriptors(path, des)
    fp = gzip.open(path, "rb")                      # Image.load_descriptors, image.py:165-167
    back = np.load(fp)
    fp.close()
    assert back.dtype == dtype and np.array_equal(back, des)
    assert np.array_equal(featcache.load_descriptors(path), des)


@pytest.mark.gpu
def test_detect_and_cache_writes_full_resolution_key_points(tmp_path):
    from conftest import load_golden
    img = load_golden("sift_reference.npz")["medium_image"]
    feat, desc = str(tmp_path / "a.feat"), str(tmp_path / "a.desc")
    kps, des = featcache.detect_and_cache(img, 0.4, feat, desc)
    back, dback = featcache.load_features(feat), featcache.load_descriptors(desc)
    assert len(back) == len(kps) == len(dback) > 1000 and dback.dtype == np.float32 and dback.shape[1] == 128
    g = load_golden("sift_reference.npz")["medium_kp"]
    assert abs(len(g) - len(back)) <= len(g) // 100
    # the first key point in cv2's order, moved back to full resolution (kp.pt / scale)
    assert abs(back[0].pt[0] - g[0, 0] / 0.4) < 0.1 and abs(back[0].pt[1] - g[0, 1] / 0.4) < 0.1
    orb = detector.ORB_create(500)
    kps, des = featcache.detect_and_cache(img, 1.0, feat, desc, det=orb)
    assert featcache.load_descriptors(desc).dtype == np.uint8 and len(featcache.load_features(feat)) == len(kps)
