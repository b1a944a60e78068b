"""BASELINE configs[0]: 16 synthetic 1920 x 1080 frames with REAL cv2.SIFT features through the matching step.
The golden (tests/golden/make_golden_config0.py) is what the UNMODIFIED reference `lib.matcher.find_matches` -- the
call of scripts/3a-matching.py:112 / process.py:291 -- stored in every image's match_list ('traditional' strategy,
GMS live, sequential work list); the frames and features are regenerated here with the generator's own code and
checked against the recorded digest, so a different OpenCV build skips instead of failing."""
import importlib.util
import os
import pickle
import types

import numpy as np
import pytest

from conftest import load_golden
from oracle import oracle

cv2 = pytest.importorskip("cv2")
_spec = importlib.util.spec_from_file_location(
    "make_golden_config0", os.path.join(os.path.dirname(__file__), "golden", "make_golden_config0.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)

_cache = {}


def _features():
    if "f" not in _cache:
        g = load_golden("reference_config0.npz")
        feats = [gen.features(f) for f in gen.frames()]
        if gen.digest(feats) != str(g["digest"]):
            pytest.skip("this OpenCV build produces different SIFT features than the one the golden was recorded with")
        _cache["f"] = (g, feats)
    return _cache["f"]


def test_config0_oracle_equals_reference_driver_on_sample_pairs():
    g, feats = _features()
    assert int(g["n"]) == 16 and len(feats) == 16 and min(len(p) for p, _ in feats) > 2500
    for i, j in ((0, 1), (7, 10), (11, 15)):
        (p1, d1), (p2, d2) = feats[i], feats[j]
        f, r = oracle.bidirectional(d1.astype(np.uint8), d2.astype(np.uint8), oracle.NORM_L2, 0.75, 270.0, threads=4,
                                    pts_q=p1, pts_t=p2, size=(gen.W, gen.H), dedupe=True)
        assert f == g["match_frame%02d_frame%02d" % (i, j)].tolist(), (i, j)
        assert r == g["match_frame%02d_frame%02d" % (j, i)].tolist(), (j, i)


@pytest.mark.gpu
def test_config0_find_matches_equals_reference_driver(tmp_path):
    """The drop-in find_matches on the 16-frame project: every match_list equals the reference's, and so do the
    meta/<name>.match pickles (image.py:219-228)."""
    from test_gpu_parity import FakeImage
    from imageanalysis_b200 import matcher
    from imageanalysis_b200.propshim import getNode
    g, feats = _features()
    matcher.gms_enabled = True
    det = getNode("/config/detector", True)
    det.setString("detector", "SIFT")
    det.setFloat("scale", gen.SCALE)
    mn = getNode("/config/matcher", True)
    mn.setFloat("match_ratio", 0.75)
    mn.setFloat("min_pairs", 25)
    cam = getNode("/config/camera", True)
    cam.setInt("width_px", gen.W)
    cam.setInt("height_px", gen.H)
    matcher.configure()
    imgs = [FakeImage("frame%02d" % i, des, pts, (0.0, 12.0 * i, -60.0)) for i, (pts, des) in enumerate(feats)]
    proj = types.SimpleNamespace(image_list=imgs, analysis_dir=str(tmp_path))
    K = np.array([[1388.0, 0, 960.0], [0, 1388.0, 540.0], [0, 0, 1]])
    matcher.find_matches(proj, K, strategy="traditional", transform="homography", sort=False, review=False)
    checked = matches = 0
    for im in imgs:
        want = {k[len("match_%s_" % im.name):]: g[k].tolist() for k in g.files if k.startswith("match_%s_" % im.name)}
        assert im.match_list == want, im.name
        assert pickle.dumps(im.match_list) == pickle.dumps(want)          # the .match file of this image
        checked += len(want)
        matches += sum(len(v) for v in want.values())
    assert checked == 108 and matches > 50000
    matcher.gms_enabled = False
