"""GMS grid filter: the oracle restatement against the reference's own
scripts/lib/archive/gms_matcher.py (goldens from tests/golden/make_golden_gms.py),
and the CUDA kernel against both."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import oracle


def _case(g, name):
    return g[name + "_pts1"], g[name + "_pts2"], g[name + "_matches"], tuple(int(v) for v in g[name + "_size"])


def test_oracle_gms_equals_reference_module():
    g = load_golden("gms_reference.npz")
    for name in g["names"]:
        pts1, pts2, matches, size = _case(g, str(name))
        mask = oracle.gms_mask(pts1, pts2, size, size, matches, with_rotation=True, with_scale=False, threshold_factor=5.0)
        assert (mask == g[str(name) + "_mask"]).all(), name
    pts1, pts2, matches, size = _case(g, "flags")
    for ws, wr in ((False, False), (True, False), (True, True)):
        mask = oracle.gms_mask(pts1, pts2, size, size, matches, with_rotation=wr, with_scale=ws, threshold_factor=5.0)
        assert (mask == g["flags_s%d_r%d_mask" % (ws, wr)]).all(), (ws, wr)


def test_oracle_gms_edge_points_are_skipped_not_wrapped():
    """Points in the last half cell have no cell in the shifted grids (index -1).  The C++ original skips such a
    match for that grid; the archive Python would index mCellPairs[-1].  The oracle follows the C++ rule: moving an
    accepted match into the last half cell can only remove it from the shifted grids, never alias it to cell 399."""
    rng = np.random.default_rng(3)
    n = 600
    p1 = rng.uniform(0.05, 0.9, (n, 2))
    p2 = p1 + rng.normal(0, 0.002, (n, 2))
    p1[:40, 0] = rng.uniform(0.976, 0.999, 40)          # x in the last half cell
    p2[:40] = p1[:40]
    size = (2000, 1000)
    pts1 = (p1 * size).astype(np.float32)
    pts2 = (p2 * size).astype(np.float32)
    m = np.stack([np.arange(n), np.arange(n)], 1)
    mask = oracle.gms_mask(pts1, pts2, size, size, m)
    assert mask[40:].mean() > 0.9
    assert mask.dtype == bool and len(mask) == n
