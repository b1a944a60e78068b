"""SIFT detect + describe (reference scripts/lib/image.py:236-237, :324: cv2.SIFT_create().detectAndCompute) --
the CPU restatement (oracle/sift.py) against goldens recorded from live cv2 (tests/golden/make_golden_sift.py), and
the CUDA implementation (csrc/sift.cu through iam_sift_detect) against both.

SIFT is floating point end to end, so the bar is a tolerance, stated here:
  * key points: a key point of cv2 counts as reproduced when one of ours lies within 0.02 px, 0.02 in size and 0.5
    degree in orientation; at least 99 % of cv2's must be reproduced and the total count may differ by at most 1 %.
    (The separable float blur sums its taps in another order than OpenCV's SIMD filter, single pixels of the pyramid
    differ by <= 1e-4 grey levels, and an extremum / peak / Newton-step decision that sits on its boundary flips.)
  * on reproduced key points: layer and octave identical, response within 1e-4, descriptors within +-1 per byte
    (at most 2 % of descriptors may hold a byte that is off by 2, at most 0.1 % anything larger: an orientation that
    moved by a fraction of a degree inside the matching tolerance).
  * GPU against the restatement (same tap order): at least 99.5 % reproduced with the same descriptor bar; the only
    differences left are exp/pow library roundings."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import sift as S


def _check(kp, octv, des, kp_ref, octv_ref, des_ref, what, min_frac=0.99):
    assert des.shape == (len(kp), 128) and des.dtype == np.uint8 and octv.shape == (len(kp),)
    assert abs(len(kp) - len(kp_ref)) <= max(2, len(kp_ref) // 100), (what, len(kp), len(kp_ref))
    m = S.match_keypoints(kp_ref, kp)
    ok = m >= 0
    assert ok.sum() >= len(kp_ref) - max(1, int(np.ceil((1 - min_frac) * len(kp_ref)))), (what, int(ok.sum()), len(kp_ref))
    assert len(set(m[ok].tolist())) == int(ok.sum()), (what, "one key point matched twice")
    assert np.array_equal(octv_ref[ok] & 0xFFFF, octv[m[ok]] & 0xFFFF), what
    assert np.abs((octv_ref[ok] >> 16) - (octv[m[ok]] >> 16)).max() <= 1, what
    assert np.abs(kp_ref[ok, 4] - kp[m[ok], 4]).max() < 1e-4, what
    dd = np.abs(des_ref[ok].astype(int) - des[m[ok]].astype(int)).max(axis=1)
    assert (dd > 2).mean() <= 0.001 and (dd > 1).mean() <= 0.02, (what, int(dd.max()), float((dd > 1).mean()))
    return float(ok.mean()), float((dd == 0).mean())


def test_oracle_primitives():
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (9, 7)).astype(np.float32)
    up = S.np_up(img)
    assert up.shape == (18, 14) and up[0, 0] == img[0, 0] and up[-1, -1] == img[-1, -1]
    assert up[0, 1] == np.float32(img[0, 0] * np.float32(0.75) + img[0, 1] * np.float32(0.25))
    assert np.array_equal(S.np_half(up), up[::2, ::2])
    k = S.gauss_kernel(1.6)
    assert len(k) == 15 and abs(float(k.sum()) - 1) < 1e-6 and np.array_equal(k, k[::-1])
    flat = np.full((12, 10), 37, np.float32)
    assert np.abs(S.np_blur(flat, 2.0) - 37).max() < 1e-4       # reflect-101 borders keep a constant image constant
    a = S.fast_atan2(np.array([0, 1, 0, -1, 1], np.float32), np.array([1, 0, -1, 0, 1], np.float32))
    assert np.abs(a - np.array([0, 90, 180, 270, 45])).max() < 0.02


@pytest.mark.parametrize("name", ["small", "tiny"])
def test_oracle_sift_equals_cv2_small(name):
    g = load_golden("sift_reference.npz")
    kp, octv, des = S.detect_arrays(g[name + "_image"])
    frac, same = _check(kp, octv, des, g[name + "_kp"], g[name + "_octave"], g[name + "_des"], name)
    assert frac == 1.0        # recorded by make_golden_sift.py: all 254 / 39 reproduced


def test_golden_is_well_formed():
    g = load_golden("sift_reference.npz")
    for name in ("small", "medium", "large", "odd", "tiny"):
        kp, octv, des = g[name + "_kp"], g[name + "_octave"], g[name + "_des"]
        assert len(kp) == len(octv) == len(des) > 30 and des.dtype == np.uint8
        key = list(zip(kp[:, 0].tolist(), kp[:, 1].tolist(), (-kp[:, 2]).tolist(), kp[:, 3].tolist()))
        assert key == sorted(key)                  # cv2's order: x, y, size descending, angle
        layer = (octv >> 8) & 255
        assert layer.min() >= 1 and layer.max() <= 3


def test_make_detector_dispatch():
    from imageanalysis_b200 import _capi, detector
    assert isinstance(detector.make_detector("SIFT"), detector.SIFT)
    orb = detector.make_detector("ORB", 1234)
    assert isinstance(orb, detector.ORB) and orb.nfeatures == 1234
    with pytest.raises(_capi.IamError):
        detector.make_detector("SURF")


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["small", "medium", "large", "odd", "tiny"])
def test_gpu_sift_equals_cv2(name):
    from imageanalysis_b200 import detector
    g = load_golden("sift_reference.npz")
    r = detector.sift_detect_and_compute(g[name + "_image"])
    kp = np.column_stack([r["pt"], r["size"], r["angle"], r["response"]]).astype(np.float32)
    _check(kp, r["octave"], r["des"], g[name + "_kp"], g[name + "_octave"], g[name + "_des"], name)
    key = list(zip(kp[:, 0].tolist(), kp[:, 1].tolist(), (-kp[:, 2]).tolist(), kp[:, 3].tolist()))
    assert key == sorted(key)


@pytest.mark.gpu
def test_gpu_sift_equals_oracle():
    from imageanalysis_b200 import detector
    g = load_golden("sift_reference.npz")
    img = g["small_image"]
    r = detector.sift_detect_and_compute(img)
    kp = np.column_stack([r["pt"], r["size"], r["angle"], r["response"]]).astype(np.float32)
    ko, oo, do = S.detect_arrays(img)
    _check(kp, r["octave"], r["des"], ko, oo, do, "vs restatement", min_frac=0.995)


@pytest.mark.gpu
def test_gpu_sift_degenerate_inputs():
    from imageanalysis_b200 import _capi, detector
    eng = detector._eng()
    for shape in ((2, 2), (5, 300), (300, 5), (11, 11), (16, 16)):          # too small for any extremum: empty result, no error
        kp, octv, des = eng.sift_detect(np.random.default_rng(1).integers(0, 256, shape).astype(np.uint8))
        assert kp.shape == (0, 5) and octv.shape == (0,) and des.shape == (0, 128)
    with pytest.raises(_capi.IamError):
        eng.sift_detect(np.zeros((1, 50), np.uint8))
    with pytest.raises(_capi.IamError):
        eng.sift_detect(np.zeros((4, 4, 3), np.uint8))
    g = load_golden("sift_reference.npz")
    with pytest.raises(_capi.IamError, match="key points"):                  # the caller's buffers are too small: loud, not truncated
        eng.sift_detect(g["medium_image"], max_out=100)
    kp, _, _ = eng.sift_detect(g["medium_image"])                            # the context is usable afterwards
    assert abs(len(kp) - len(g["medium_kp"])) <= 15


@pytest.mark.gpu
def test_gpu_sift_many_frames_in_flight_equal_one_by_one():
    from imageanalysis_b200 import detector
    g = load_golden("sift_reference.npz")
    frames = [g["medium_image"], g["large_image"], g["small_image"], g["large_image"][::-1].copy(), g["medium_image"].T.copy()]
    one = [detector.sift_detect_and_compute(f) for f in frames]
    many = detector.sift_detect_many(frames, workers=3)
    assert len(many) == len(one)
    for a, b in zip(one, many):
        assert all(np.array_equal(a[k], b[k]) for k in ("pt", "size", "angle", "response", "octave", "des"))


@pytest.mark.gpu
def test_gpu_sift_cv2_style_api_and_matching():
    """SIFT_create().detectAndCompute(img, None) -> (key points, float32 [N, 128]) and the descriptors feed the matcher:
    a shifted copy of the image matches back with the shift."""
    from imageanalysis_b200 import detector
    g = load_golden("sift_reference.npz")
    img = g["medium_image"]
    kps, des = detector.SIFT_create().detectAndCompute(img, None)
    assert des.dtype == np.float32 and des.shape == (len(kps), 128) and len(kps) > 1000
    k = kps[0]
    assert hasattr(k, "pt") and hasattr(k, "size") and hasattr(k, "angle") and hasattr(k, "response") and hasattr(k, "octave")
    colour = np.stack([img] * 3, 2)
    kps2, des2 = detector.SIFT_create().detectAndCompute(colour, None)
    assert len(kps2) == len(kps) and (des2 == des).all()
    flat = np.full((64, 64), 128, np.uint8)
    kps3, des3 = detector.SIFT_create().detectAndCompute(flat, None)
    assert len(kps3) == 0 and des3 is None
    # same scene seen 16 px to the right: interior features reappear with x + 16 and the same descriptor
    a, b = img[:, 16:], img[:, :-16]
    ra, rb = detector.sift_detect_and_compute(a), detector.sift_detect_and_compute(b)
    from imageanalysis_b200 import _capi
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    par = _capi.Engine.make_params(match_ratio=0.7, reduce_mode=_capi.REDUCE_LOWE, cap=4000, min_pairs=0, cross_check=True)
    table, count = eng.match_images([0, 1], [ra["des"], rb["des"]], [[0, 1]], par)
    m = table[0, :count[0]]
    assert len(m) > 300
    dx = rb["pt"][m[:, 1], 0] - ra["pt"][m[:, 0], 0]
    dy = rb["pt"][m[:, 1], 1] - ra["pt"][m[:, 0], 1]
    assert np.median(np.abs(dx - 16)) < 0.05 and np.median(np.abs(dy)) < 0.05
