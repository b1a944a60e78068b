"""Host-side float32 -> uint8 narrowing used by iam_match_images for transport (csrc/host_narrow.cpp).
No GPU needed.  The reference keeps SIFT descriptors as float32 arrays of integers (image.py:160-180)."""
import numpy as np
import pytest

from imageanalysis_b200 import _capi


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 128, 5000 * 128 + 7])
def test_narrow_exact(n):
    rng = np.random.default_rng(n)
    a = rng.integers(0, 256, size=n).astype(np.float32)
    out, ok = _capi.narrow_host(a)
    assert ok
    assert np.array_equal(out, a.astype(np.uint8))


@pytest.mark.parametrize("bad", [0.5, 255.5, 256.0, -1.0, 1e9, -1e9, np.nan, np.inf, 3.0000002])
@pytest.mark.parametrize("pos", [0, 17, 63, 64, 99])
def test_narrow_rejects(bad, pos):
    a = np.arange(100, dtype=np.float32)
    a[pos] = bad
    _, ok = _capi.narrow_host(a)
    assert not ok


def test_narrow_negative_zero_and_edges():
    a = np.array([-0.0, 0.0, 255.0, 1.0] * 16, dtype=np.float32)
    out, ok = _capi.narrow_host(a)
    assert ok and np.array_equal(out, np.array([0, 0, 255, 1] * 16, dtype=np.uint8))
