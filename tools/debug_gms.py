#!/usr/bin/env python3
"""GPU debug aid: the match pipeline with the GMS stage, stage by stage against the oracle."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden  # noqa: E402
from imageanalysis_b200 import _capi, matcher  # noqa: E402
from oracle import oracle  # noqa: E402

g = load_golden("reference_gms_pipeline.npz")
size = (5472, 3648)
eng = _capi.Engine(_capi.NORM_L2, 128, 0)
for i in range(3):
    eng.upload(i, g["des%d" % i])
    eng.upload_keypoints(i, g["pts%d" % i])
    eng.upload_keypoint_keys(i, matcher.keypoint_keys([types.SimpleNamespace(pt=(float(x), float(y))) for x, y in g["pts%d" % i]]))
for a, b in ((0, 1), (1, 0), (1, 2)):
    kw = dict(norm=oracle.NORM_L2, match_ratio=0.75, max_distance=270.0, threads=8)
    for tag, okw, pkw in (("reduce only", dict(), dict()),
                          ("+gms", dict(size=size), dict(gms=True, size=size)),
                          ("+gms+dedupe", dict(size=size, dedupe=True), dict(gms=True, size=size, dedupe=True)),
                          ("+dedupe", dict(dedupe=True), dict(dedupe=True))):
        want = oracle.basic_pair(g["des%d" % a], g["des%d" % b], pts_q=g["pts%d" % a], pts_t=g["pts%d" % b], **kw, **okw)
        prm = _capi.Engine.make_params(cross_check=False, **pkw)
        table, count = eng.match_pairs([(a, b)], prm)
        got = table[0, :count[0]].tolist()
        same = got == want
        print("pair", (a, b), tag, "gpu", len(got), "oracle", len(want), "equal", same)
        if not same:
            sg, sw = {tuple(x) for x in got}, {tuple(x) for x in want}
            print("   only gpu:", sorted(sg - sw)[:8], " only oracle:", sorted(sw - sg)[:8])
            if tag == "+gms":
                red = oracle.basic_pair(g["des%d" % a], g["des%d" % b], **kw)
                m2 = eng.gms_filter(g["pts%d" % a], g["pts%d" % b], np.int32(red), size)
                print("   standalone gms on the oracle's reduced list:", int(m2.sum()), "kept; equals oracle:",
                      [p for p, k in zip(red, m2) if k] == want)
