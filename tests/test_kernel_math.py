"""The integer / bit-pattern identities the packed-key epilogue of csrc/knn_umma.cu relies on, restated in numpy and
checked against plain sorting.  No GPU needed: these guard the arithmetic, the kernel itself is
checked against the oracle in test_gpu_parity.py."""
import numpy as np
from hypothesis import given, settings, strategies as st


def packed_top2_i8(acc):
    """consume32_packed: acc [32] ints < 2^24, larger = nearer.  Returns ((value, column), (value, column))."""
    j = np.arange(32)
    key = (acc.astype(np.int64) * 32 + (31 - j)).astype(np.uint32)           # unique 29-bit keys
    m1 = key.max()
    u = (key - m1).astype(np.uint32)                                          # mod 2^32: winner 0, the others huge
    m2 = (int(m1) + int(u.max())) % (1 << 32)                                 # wraps back to the runner-up's key
    return [(int(m) >> 5, 31 - (int(m) & 31)) for m in (m1, m2)]


def packed_top2_hamming(key):
    """consume32_packed_f: key [32] float32 = 32 * distance + column (exact integers), smaller = nearer."""
    key = key.astype(np.float32)
    m1 = key.min()
    t = (key - np.float32(m1 + np.float32(0.25))).astype(np.float32)         # winner -1/4, others n - 1/4, n >= 1
    bits = t.view(np.uint32)                                                  # the one negative float is the largest unsigned
    m2 = np.float32(np.array(bits.min(), np.uint32).view(np.float32)) + np.float32(m1 + np.float32(0.25))
    out = []
    for m in (m1, m2):
        q = np.array(np.float32(m) + np.float32(8388608.0), np.float32).view(np.uint32)   # key in the mantissa
        col = int(q & 31)
        out.append((int((float(m) - col) / 32), col))
    return out


@settings(max_examples=300, deadline=None)
@given(st.lists(st.integers(0, (1 << 24) - 1), min_size=32, max_size=32), st.integers(0, 5))
def test_packed_top2_byte_layout(vals, dup):
    acc = np.array(vals, np.int64)
    if dup:                                   # force ties: the lower column must win
        acc[dup * 5] = acc.max()
        acc[31 - dup] = acc.max()
    order = sorted(range(32), key=lambda c: (-acc[c], c))
    want = [(int(acc[c]), c) for c in order[:2]]
    assert packed_top2_i8(acc) == want


@settings(max_examples=300, deadline=None)
@given(st.lists(st.integers(0, 256), min_size=32, max_size=32), st.integers(0, 5), st.booleans())
def test_packed_top2_hamming_keys(dists, dup, padding):
    d = np.array(dists, np.int64)
    if dup:
        d[dup * 3] = d.min()
        d[31 - dup] = d.min()
    if padding:
        d[20:] = 2048                         # padding train rows: key 65536 + column
    key = (32 * d + np.arange(32)).astype(np.float32)
    order = sorted(range(32), key=lambda c: (d[c], c))
    want = [(int(d[c]), c) for c in order[:2]]
    assert packed_top2_hamming(key) == want


@settings(max_examples=500, deadline=None)
@given(st.lists(st.tuples(st.integers(-1, 1 << 24), st.integers(-1, 1 << 24)), min_size=3, max_size=3))
def test_row_exact_second_best(parts):
    """Each of the three threads of a row publishes (best, second best), best >= second; the row's second best is
    max(second largest of the bests, largest of the second bests)."""
    hl = [(max(a, b), min(a, b)) for a, b in parts]
    h = [x[0] for x in hl]
    lo = [x[1] for x in hl]
    second_h = max(min(h[0], h[1]), min(h[0], h[2]), min(h[1], h[2]))
    got = max(second_h, max(lo))
    want = sorted(h + lo, reverse=True)[1]
    assert got == want


def test_hamming_operand_layout_gives_keys():
    """convert_hamming_kernel: A = [bits | 32, 256, p0, 2*p1, 256, 1, 1], B = [-64 * bits | p0, 2*p1, 32, 256, pad, j & 15,
    j & 16] with popcount = p0 + 16 * p1: the dot product is 32 * Hamming distance + (train row mod 32)."""
    rng = np.random.default_rng(3)
    q = rng.integers(0, 256, size=(40, 32), dtype=np.uint8)
    t = rng.integers(0, 256, size=(70, 32), dtype=np.uint8)
    q[0] = 255
    t[1] = 255
    t[2] = 0
    qb, tb = np.unpackbits(q, axis=1).astype(np.int64), np.unpackbits(t, axis=1).astype(np.int64)

    def aug(bits, role, rows):
        pc = bits.sum(1)
        p0, p1 = pc & 15, pc >> 4
        r = np.arange(rows)
        if role == "a":
            return np.stack([np.full(rows, 32), np.full(rows, 256), p0, 2 * p1, np.full(rows, 256), np.ones(rows, int), np.ones(rows, int)], 1)
        return np.stack([p0, 2 * p1, np.full(rows, 32), np.full(rows, 256), np.zeros(rows, int), r & 15, r & 16], 1)
    A = np.concatenate([qb, aug(qb, "a", 40)], 1)
    B = np.concatenate([-64 * tb, aug(tb, "b", 70)], 1)
    acc = A @ B.T
    ham = (qb[:, None, :] != tb[None, :, :]).sum(-1)
    assert (acc == 32 * ham + (np.arange(70) % 32)[None, :]).all()
    assert acc.max() < (1 << 14) and np.abs(-64 * (qb @ tb.T)).max() <= (1 << 14)   # exact in any fp32 accumulation order
