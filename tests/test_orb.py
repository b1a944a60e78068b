"""ORB detect + describe (reference scripts/lib/image.py:243-245, :324: cv2.ORB_create(n).detectAndCompute) --
the CPU restatement (oracle/orb.py) against goldens recorded from live cv2 (tests/golden/make_golden_orb.py), and
the CUDA implementation (csrc/orb.cu through iam_orb_detect) against both.

Bar: key point SETS (level coordinates, octave) identical; Harris responses bit exact; orientation within 1e-3 degree
(OpenCV's fastAtan2 is a float polynomial whose low bits depend on the compiler's evaluation order); descriptors: at
most 1 differing bit per 1000 descriptors (a test pixel lands on the other side of a .5 rounding of the float blur or
of the rotated sampling position)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import orb as O


def _as_dict(pt, octave, angle, response, size, des):
    return {(round(float(p[0]), 2), round(float(p[1]), 2), int(o)): (float(a), float(r), float(s), bytes(d))
            for p, o, a, r, s, d in zip(pt, octave, angle, response, size, des)}


def _golden(g, tag):
    return _as_dict(g[tag + "_pt"], g[tag + "_octave"], g[tag + "_angle"], g[tag + "_response"], g[tag + "_size"], g[tag + "_des"])


def _compare(got, want, what):
    assert set(got) == set(want), (what, len(set(got) ^ set(want)))
    bits = 0
    for k in want:
        a0, r0, s0, d0 = want[k]
        a1, r1, s1, d1 = got[k]
        da = abs(a0 - a1)
        assert min(da, 360 - da) < 1e-3, (what, k, a0, a1)
        assert r0 == r1 and s0 == s1, (what, k)
        bits += int(np.unpackbits(np.frombuffer(d0, np.uint8) ^ np.frombuffer(d1, np.uint8)).sum())
    assert bits <= max(1, len(want) // 1000), (what, bits)
    return bits


def test_oracle_stages_equal_cv2_goldens():
    g = load_golden("orb_reference.npz")
    for name in ("texture", "blocks"):
        img = g[name + "_img"]
        xs, ys, sc = O.fast_detect(img)
        assert sorted(zip(xs.tolist(), ys.tolist(), sc.astype(int).tolist())) == [tuple(r) for r in g[name + "_fast"].tolist()]
    assert O.features_per_level(500) == [109, 90, 75, 63, 52, 44, 36, 31] and sum(O.features_per_level(20000)) == 20000
    assert O.umax_table()[:16] == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


@pytest.mark.parametrize("name,n", [("texture", 500), ("texture", 2000), ("blocks", 500), ("blocks", 2000)])
def test_oracle_orb_equals_cv2(name, n):
    g = load_golden("orb_reference.npz")
    r = O.detect_and_compute(g[name + "_img"], n)
    got = _as_dict(r["pt"], r["octave"], r["angle"], r["response"], r["size"], r["des"])
    _compare(got, _golden(g, "%s_%d" % (name, n)), (name, n))


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name,n", [("texture", 500), ("texture", 2000), ("blocks", 500), ("blocks", 2000)])
def test_gpu_orb_equals_cv2_and_oracle(name, n):
    from imageanalysis_b200 import detector
    g = load_golden("orb_reference.npz")
    img = g[name + "_img"]
    r = detector.orb_detect_and_compute(img, n)
    got = _as_dict(r["pt"], r["octave"], r["angle"], r["response"], r["size"], r["des"])
    _compare(got, _golden(g, "%s_%d" % (name, n)), (name, n, "vs cv2"))
    o = O.detect_and_compute(img, n)
    _compare(got, _as_dict(o["pt"], o["octave"], o["angle"], o["response"], o["size"], o["des"]), (name, n, "vs oracle"))


@pytest.mark.gpu
def test_gpu_orb_cv2_style_api():
    """ORB_create(n).detectAndCompute(img, None) -> (key points with .pt/.size/.angle/.response/.octave, uint8 [N, 32])."""
    from imageanalysis_b200 import detector
    g = load_golden("orb_reference.npz")
    kps, des = detector.ORB_create(500).detectAndCompute(g["texture_img"], None)
    assert len(kps) == 500 and des.shape == (500, 32) and des.dtype == np.uint8
    k = kps[0]
    assert hasattr(k, "pt") and hasattr(k, "size") and hasattr(k, "angle") and hasattr(k, "response") and hasattr(k, "octave")
    colour = np.stack([g["texture_img"]] * 3, 2)       # BGR input is converted to grey like cv2 does
    kps2, des2 = detector.ORB_create(500).detectAndCompute(colour, None)
    assert len(kps2) == 500 and (des2 == des).all()
    tiny = np.zeros((40, 40), np.uint8)
    kps3, des3 = detector.ORB_create(500).detectAndCompute(tiny, None)
    assert len(kps3) == 0 and des3 is None
