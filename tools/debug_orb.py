#!/usr/bin/env python3
"""Debug aid: GPU ORB (iam_orb_detect) against the CPU restatement, per pyramid level."""
import collections
import sys

import numpy as np

sys.path.insert(0, ".")
from imageanalysis_b200 import detector  # noqa: E402
from oracle import orb as O  # noqa: E402

g = np.load("tests/golden/orb_reference.npz")
img = g["texture_img"]
n = 500
r = detector.orb_detect_and_compute(img, n)
o = O.detect_and_compute(img, n)
levels, scales = O.build_pyramid(img)


def lv(d):
    out = collections.defaultdict(dict)
    for p, oc, a, rs, de in zip(d["pt"], d["octave"], d["angle"], d["response"], d["des"]):
        s = scales[int(oc)]
        out[int(oc)][(int(round(float(p[0]) / s)), int(round(float(p[1]) / s)))] = (float(a), float(rs), bytes(de))
    return out


G, R = lv(r), lv(o)
for l in range(8):
    a, b = G.get(l, {}), R.get(l, {})
    common = set(a) & set(b)
    same_resp = sum(a[k][1] == b[k][1] for k in common)
    same_des = sum(a[k][2] == b[k][2] for k in common)
    dang = max([min(abs(a[k][0] - b[k][0]), 360 - abs(a[k][0] - b[k][0])) for k in common] or [0])
    print("level %d: gpu %4d oracle %4d common %4d  same response %4d  same descriptor %4d  max angle diff %.4f" % (
        l, len(a), len(b), len(common), same_resp, same_des, dang))
    og = sorted(set(a) - set(b))[:4]
    oo = sorted(set(b) - set(a))[:4]
    if og or oo:
        print("   only gpu:", [(k, round(a[k][1], 8)) for k in og], " only oracle:", [(k, round(b[k][1], 8)) for k in oo])
        if common:
            k = sorted(common)[0]
            print("   a common point", k, "gpu resp %.9g oracle resp %.9g" % (a[k][1], b[k][1]))
