"""GPU parity tests: the CUDA path, called through the C-ABI (ctypes), against
the oracle and the committed golden vectors.  Integer paths are compared
bit-for-bit (indices AND float32 distances)."""
import os
import types

import numpy as np
import pytest

from conftest import load_golden
from imageanalysis_b200 import _capi, synth
from oracle import oracle

pytestmark = pytest.mark.gpu

# UMMA: the tcgen05 kernel with the operand kind the library picks (byte operands / kind::i8 for integer-valued L2,
# e4m3 for Hamming); UMMA_F16: the same kernel forced onto fp16 operands; SIMT: the CUDA-core cross-check engine.
ENGINES = [(_capi.ENGINE_UMMA, "umma"), (_capi.ENGINE_UMMA_F16, "umma_f16"), (_capi.ENGINE_SIMT, "simt")]
KNN_CASES = [("knn_l2_synth.npz", _capi.NORM_L2), ("knn_hamming_synth.npz", _capi.NORM_HAMMING),
             ("knn_sift_real.npz", _capi.NORM_L2), ("knn_orb_real.npz", _capi.NORM_HAMMING)]


def run_knn(norm, q, t, k, engine, reverse=True):
    eng = _capi.Engine(norm, q.shape[1], 0)
    eng.set_engine(engine)
    eng.upload(0, q)
    eng.upload(1, t)
    n = max(q.shape[0], t.shape[0])
    out = eng.knn_pairs([(0, 1)], k, n, reverse=reverse)
    tm = eng.timing()
    eng.close()
    assert tm.engine_used == (_capi.ENGINE_SIMT if engine == _capi.ENGINE_SIMT else _capi.ENGINE_UMMA)
    if engine == _capi.ENGINE_UMMA_F16 and norm == _capi.NORM_L2:
        assert tm.mma_kind == _capi.KIND_F16
    return out


@pytest.mark.parametrize("engine,ename", ENGINES)
@pytest.mark.parametrize("name,norm", KNN_CASES)
@pytest.mark.parametrize("k", [1, 2, 3])
def test_knn_equals_cv2_golden(engine, ename, name, norm, k):
    g = load_golden(name)
    q, t = g["q"], g["t"]
    idx, dist, ridx, rdist = run_knn(norm, q, t, k, engine)
    assert (idx[0, :len(q)] == g["idx"][:, :k]).all()
    assert (dist[0, :len(q)] == g["dist"][:, :k]).all()
    assert (ridx[0, :len(t)] == g["ridx"][:, :k]).all()
    assert (rdist[0, :len(t)] == g["rdist"][:, :k]).all()


@pytest.mark.parametrize("engine,ename", ENGINES)
def test_float32_sift_input_is_exact_path(engine, ename):
    g = load_golden("knn_sift_real.npz")
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    eng.set_engine(engine)
    eng.upload(0, g["q"].astype(np.float32))     # what cv2.SIFT hands the reference (image.py:324)
    eng.upload(1, g["t"].astype(np.float32))
    assert eng.descriptors_exact(0) == 1 and eng.num_descriptors(1) == len(g["t"])
    idx, dist, _, _ = eng.knn_pairs([(0, 1)], 2, len(g["q"]), reverse=False)
    assert (idx[0] == g["idx"][:, :2]).all() and (dist[0] == g["dist"][:, :2]).all()


@pytest.mark.parametrize("norm,gen,nb", [(_capi.NORM_L2, synth.sift_like, 128), (_capi.NORM_HAMMING, synth.orb_like, 32)])
@pytest.mark.parametrize("nq,nt", [(1, 5), (2, 2), (127, 129), (128, 128), (129, 127), (255, 257), (256, 256),
                                   (300, 1), (513, 1025), (1000, 3), (2000, 1777)])
def test_knn_ragged_sizes(norm, gen, nb, nq, nt):
    q, t = gen(nq, seed=nq), gen(nt, seed=1000 + nt)
    if nt > 4 and nq > 4:
        t[nt // 2] = q[1]
        t[nt - 1] = q[1]                      # duplicate at the very last valid column of the last tile
        t[0] = q[nq - 1]
    for engine, _ in ENGINES:
        idx, dist, ridx, rdist = run_knn(norm, q, t, 3, engine)
        oi, od = oracle.knn(q, t, 3, norm, threads=4)
        ri, rd = oracle.knn(t, q, 3, norm, threads=4)
        assert (idx[0, :nq] == oi).all() and (dist[0, :nq] == od).all()
        assert (ridx[0, :nt] == ri).all() and (rdist[0, :nt] == rd).all()


def test_hamming_extreme_popcounts_stay_exact():
    """The Hamming operand layout makes the accumulator a key, 32 * distance + column (convert_hamming_kernel), out of
    partial sums up to 32 * 512 = 16384 in magnitude: saturated, empty and nearly-saturated descriptors must still
    give exact distances and cv2's tie order, for every k."""
    rng = np.random.default_rng(7)
    q = np.zeros((200, 32), np.uint8)
    q[::3] = 255                                   # all ones
    q[1::3] = rng.integers(0, 256, size=(67, 32), dtype=np.uint8) | rng.integers(0, 256, size=(67, 32), dtype=np.uint8) | 0xF0
    t = np.full((333, 32), 255, np.uint8)
    t[::4] = 0
    t[1::4, :5] = rng.integers(0, 256, size=(83, 5), dtype=np.uint8)
    t[2::4] = rng.integers(0, 256, size=(83, 32), dtype=np.uint8) | 0x0F
    for k in (1, 2, 3):
        for engine, _ in ENGINES:
            idx, dist, ridx, rdist = run_knn(_capi.NORM_HAMMING, q, t, k, engine)
            oi, od = oracle.knn(q, t, k, _capi.NORM_HAMMING, threads=4)
            ri, rd = oracle.knn(t, q, k, _capi.NORM_HAMMING, threads=4)
            assert (idx[0, :200] == oi).all() and (dist[0, :200] == od).all()
            assert (ridx[0, :333] == ri).all() and (rdist[0, :333] == rd).all()


def test_extreme_values_stay_exact():
    """all-255 vs all-0 rows: d^2 = 128*255^2 = 8 323 200, the largest value the
    fp32 accumulator must hold exactly."""
    q = np.zeros((130, 128), np.uint8)
    q[::2] = 255
    q[5, :64] = 255
    t = np.zeros((140, 128), np.uint8)
    t[1::2] = 255
    t[7, 64:] = 254
    for engine, _ in ENGINES:
        idx, dist, _, _ = run_knn(_capi.NORM_L2, q, t, 3, engine, reverse=False)
        oi, od = oracle.knn(q, t, 3, _capi.NORM_L2)
        assert (idx[0, :130] == oi).all() and (dist[0, :130] == od).all()
    # the largest representable distance: all-0 query against all-255 train rows only
    t2 = np.full((3, 128), 255, np.uint8)
    q2 = np.zeros((5, 128), np.uint8)
    for engine, _ in ENGINES:
        idx, dist, _, _ = run_knn(_capi.NORM_L2, q2, t2, 2, engine, reverse=False)
        assert (idx[0, :5] == [0, 1]).all() and (dist[0, :5] == np.float32(np.sqrt(128 * 255.0 ** 2))).all()


@pytest.mark.parametrize("norm,gen,nb", [(_capi.NORM_L2, synth.sift_like, 128), (_capi.NORM_HAMMING, synth.orb_like, 32)])
def test_full_size_5000_umma_equals_simt_and_oracle_rows(norm, gen, nb):
    """BASELINE config 2/3 shape (5000 x 5000).  Full-matrix check: the tcgen05
    engine against the independent CUDA-core engine; plus a 300-row slice
    against the CPU oracle."""
    rng = np.random.default_rng(5)
    q, t = gen(5000, seed=71), gen(5000, seed=72)
    src = rng.permutation(5000)[:2000]
    if norm == _capi.NORM_L2:
        q[:2000] = np.clip(t[src].astype(np.int32) + rng.integers(-3, 4, (2000, 128)), 0, 255)
    else:
        q[:2000] = synth.flip_bits(t[src], 20, rng)
    a = run_knn(norm, q, t, 2, _capi.ENGINE_UMMA)
    b = run_knn(norm, q, t, 2, _capi.ENGINE_SIMT)
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)
    rows = rng.permutation(5000)[:300]
    oi, od = oracle.knn(q[rows], t, 2, norm, threads=8)
    assert (a[0][0, rows] == oi).all() and (a[1][0, rows] == od).all()


def test_operand_kind_selection():
    """Integer-valued L2 descriptors run on byte operands (kind::i8); rows whose squared norm exceeds the byte
    layout's capacity, non-integer descriptors and forced fp16 fall back to fp16 operands; Hamming is e4m3."""
    def kind_of(norm, q, t, engine=_capi.ENGINE_AUTO):
        eng = _capi.Engine(norm, q.shape[1], 0)
        eng.set_engine(engine)
        eng.upload(0, q)
        eng.upload(1, t)
        idx, dist, _, _ = eng.knn_pairs([(0, 1)], 2, len(q), reverse=False)
        k = eng.timing().mma_kind
        eng.close()
        return k, idx, dist
    q, t = synth.sift_like(300, seed=1), synth.sift_like(400, seed=2)
    k, idx, dist = kind_of(_capi.NORM_L2, q, t)
    assert k == _capi.KIND_I8
    k2, idx2, dist2 = kind_of(_capi.NORM_L2, q.astype(np.float32), t.astype(np.float32))
    assert k2 == _capi.KIND_I8 and (idx2 == idx).all() and (dist2 == dist).all()
    k3, idx3, dist3 = kind_of(_capi.NORM_L2, q, t, _capi.ENGINE_UMMA_F16)
    assert k3 == _capi.KIND_F16 and (idx3 == idx).all() and (dist3 == dist).all()
    t_big = t.copy()
    t_big[7] = 255                                  # squared norm 8 323 200 > capacity 4 031 547
    k4, idx4, dist4 = kind_of(_capi.NORM_L2, q, t_big)
    oi, od = oracle.knn(q, t_big, 2, oracle.NORM_L2, threads=4)
    assert k4 == _capi.KIND_F16 and (idx4[0] == oi).all() and (dist4[0] == od).all()
    k5, _, _ = kind_of(_capi.NORM_L2, q.astype(np.float32) + 0.25, t.astype(np.float32))
    assert k5 == _capi.KIND_F16
    k6, _, _ = kind_of(_capi.NORM_HAMMING, synth.orb_like(300, seed=1), synth.orb_like(300, seed=2))
    assert k6 == _capi.KIND_F8


@pytest.mark.parametrize("k", [1, 2, 3])
def test_byte_layout_norm_parity_and_ties(k):
    """The byte layout orders train rows by the parity of their squared norm and compares q.t - floor(|t|^2/2):
    small-valued descriptors give many exactly equal distances within and across the two parity classes, the
    largest eligible norms sit at the capacity limit -- results must still be cv2's (distance, then lowest index)."""
    rng = np.random.default_rng(44)
    q = rng.integers(0, 3, (700, 128)).astype(np.uint8)
    t = rng.integers(0, 3, (900, 128)).astype(np.uint8)
    t[100:160] = q[:60]                      # exact duplicates (distance 0) ...
    t[500:560] = q[:60]                      # ... twice: ties on the best AND second best
    t[300] = t[301] = t[0]
    q[650:] = 0                              # all-zero queries: distance = |t|^2, every parity mix
    big = np.zeros((4, 128), np.uint8)
    big[:, :61] = 255
    big[:, 61:67] = [254, 22, 4, 2, 1, 1]    # squared norm 4 031 547: the largest the byte layout holds (odd)
    big[1, 66] = 0                           # ... and its even neighbour 4 031 546
    big[2, :] = big[2, ::-1].copy()
    assert int((big[0].astype(np.int64) ** 2).sum()) == 4031547
    t[600:604] = big
    q[600:604] = big
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    eng.upload(0, q)
    eng.upload(1, t)
    idx, dist, ridx, rdist = eng.knn_pairs([(0, 1)], k, 900)
    assert eng.timing().mma_kind == _capi.KIND_I8
    eng.close()
    oi, od = oracle.knn(q, t, k, oracle.NORM_L2, threads=4)
    ri, rd = oracle.knn(t, q, k, oracle.NORM_L2, threads=4)
    assert (idx[0, :700] == oi).all() and (dist[0, :700] == od).all()
    assert (ridx[0, :900] == ri).all() and (rdist[0, :900] == rd).all()


def test_match_images_falls_back_to_fp16_operands_once():
    """iam_match_images builds the byte layout optimistically; a descriptor set it cannot hold voids the run,
    the call repeats itself on fp16 operands and the context stays there."""
    des, _, _ = synth.sift_project(6, 600, seed=33)
    des = [d.copy() for d in des]
    des[4][17] = 255                          # one row beyond the byte layout's capacity
    pairs = [(i, j) for i in range(6) for j in range(i + 1, 6)]
    prm = _capi.Engine.make_params(cross_check=True)
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    for rep in range(2):
        table, count = eng.match_images(list(range(6)), des, pairs, prm)
        assert eng.timing().mma_kind == _capi.KIND_F16
        for p, (a, b) in enumerate(pairs):
            f, _ = oracle.bidirectional(des[a], des[b], oracle.NORM_L2, 0.75, 270.0, threads=4)
            assert table[p, :count[p]].tolist() == f
    eng.close()
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)   # the same project without the offending row stays on byte operands
    des[4][17] = des[4][16]
    table, count = eng.match_images(list(range(6)), des, pairs, prm)
    assert eng.timing().mma_kind == _capi.KIND_I8
    for p, (a, b) in enumerate(pairs):
        f, _ = oracle.bidirectional(des[a], des[b], oracle.NORM_L2, 0.75, 270.0, threads=4)
        assert table[p, :count[p]].tolist() == f
    eng.close()


def test_many_pairs_and_chunking(monkeypatch):
    """Several images, all pairs, forced through multiple workspace chunks."""
    des, _, _ = synth.sift_project(6, 700, seed=2)
    pairs = [(i, j) for i in range(6) for j in range(i + 1, 6)]
    ref = {}
    for engine, _ in ENGINES:
        for chunk in ("", "1"):
            if chunk:
                monkeypatch.setenv("IAM_CHUNK_MB", chunk)
            else:
                monkeypatch.delenv("IAM_CHUNK_MB", raising=False)
            eng = _capi.Engine(_capi.NORM_L2, 128, 0)
            eng.set_engine(engine)
            for i, d in enumerate(des):
                eng.upload(i, d)
            out = eng.knn_pairs(pairs, 2, 700)
            prm = _capi.Engine.make_params()
            tab = eng.match_pairs(pairs, prm)
            eng.close()
            if not ref:
                ref["knn"], ref["tab"] = out, tab
            else:
                for x, y in zip(out, ref["knn"]):
                    assert np.array_equal(x, y, equal_nan=True)
                assert (tab[1] == ref["tab"][1]).all()
                for p in range(len(pairs)):
                    assert (tab[0][p, :tab[1][p]] == ref["tab"][0][p, :tab[1][p]]).all()
    for p, (i, j) in enumerate(pairs):
        oi, od = oracle.knn(des[i], des[j], 2, oracle.NORM_L2, threads=4)
        assert (ref["knn"][0][p] == oi).all() and (ref["knn"][1][p] == od).all()


@pytest.mark.parametrize("mode", [_capi.REDUCE_REF_METRIC, _capi.REDUCE_LOWE])
@pytest.mark.parametrize("cross", [True, False])
def test_match_tables_equal_oracle(mode, cross):
    des, _, _ = synth.sift_project(5, 1500, seed=4)
    pairs = [(0, 1), (1, 2), (0, 2), (0, 4), (3, 4)]
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    for i, d in enumerate(des):
        eng.upload(i, d)
    prm = _capi.Engine.make_params(reduce_mode=mode, cross_check=cross, min_pairs=25, cap=500)
    table, count = eng.match_pairs(pairs, prm)
    for p, (i, j) in enumerate(pairs):
        f, _ = oracle.bidirectional(des[i], des[j], oracle.NORM_L2, 0.75, 270.0, cap=500, min_pairs=25, threads=4,
                                    mode="ref_metric" if mode == _capi.REDUCE_REF_METRIC else "lowe",
                                    cross_check=cross)
        assert table[p, :count[p]].tolist() == f
    assert count[0] > 100 and count[3] < count[0] // 4   # neighbours match strongly, frames 4 apart barely


def test_compact_tables_equal_padded_tables():
    """iam_fetch_packed_tables / iam_pack_tables_device: the CSR form holds exactly the valid rows of the padded tables,
    pair by pair; a too-small caller array is refused loudly with the size needed."""
    des, _, _ = synth.sift_project(6, 900, seed=9)
    pairs = [(0, 1), (1, 2), (0, 5), (2, 3), (3, 4), (4, 5), (0, 3)]
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    for i, d in enumerate(des):
        eng.upload(i, d)
    prm = _capi.Engine.make_params(cross_check=True, min_pairs=25, cap=700)
    eng.match_pairs_device(np.int32(pairs), prm)
    table, count = eng.fetch_tables(len(pairs), prm.cap)
    rows, off = eng.fetch_packed_tables(len(pairs))
    assert off[0] == 0 and (np.diff(off) == count).all() and rows.shape == (int(count.sum()), 2) and count.max() > 100
    for p in range(len(pairs)):
        assert (rows[off[p]:off[p + 1]] == table[p, :count[p]]).all()
    small = np.empty((10, 2), np.int32)
    with pytest.raises(_capi.IamError, match="match rows"):
        eng.fetch_packed_tables(len(pairs), rows_out=small)
    big = np.empty((int(count.sum()) + 7, 2), np.int32)
    rows2, off2 = eng.fetch_packed_tables(len(pairs), rows_out=big, offsets_out=np.empty(len(pairs) + 1, np.int32))
    assert (rows2 == rows).all() and (off2 == off).all()
    eng.close()


def test_reduce_paths_small_sort_and_large_rank():
    """More than 8192 candidates per direction leaves the shared-memory sort for the
    rank-by-counting path; both must order exactly like Python's stable sorted()."""
    rng = np.random.default_rng(12)
    a = synth.sift_like(9000, seed=91)
    b = np.clip(a[rng.permutation(9000)].astype(np.int32) + rng.integers(-2, 3, (9000, 128)), 0, 255).astype(np.uint8)
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    eng.upload(0, a)
    eng.upload(1, b)
    for cap in (2000, 9000):
        prm = _capi.Engine.make_params(cap=cap, cross_check=False)
        table, count = eng.match_pairs([(0, 1)], prm)
        f, _ = oracle.bidirectional(a, b, oracle.NORM_L2, 0.75, 270.0, cap=cap, threads=8, cross_check=False)
        assert count[0] == len(f) == cap
        assert table[0, :count[0]].tolist() == f


def test_orb_match_tables_equal_oracle():
    des, _, _ = synth.sift_project(3, 1200, seed=6, kind="orb")
    eng = _capi.Engine(_capi.NORM_HAMMING, 32, 0)
    for i, d in enumerate(des):
        eng.upload(i, d)
    prm = _capi.Engine.make_params(max_distance=64.0)
    table, count = eng.match_pairs([(0, 1), (1, 2)], prm)
    for p, (i, j) in enumerate([(0, 1), (1, 2)]):
        f, _ = oracle.bidirectional(des[i], des[j], oracle.NORM_HAMMING, 0.75, 64.0, threads=4)
        assert table[p, :count[p]].tolist() == f
    assert count.min() > 50


def test_dedupe_and_cross_check_equal_reference_module():
    """Device filter_duplicates + cross-check against what the reference's own
    lib.matcher produced (golden reference_reductions.npz)."""
    from imageanalysis_b200 import matcher
    g = load_golden("reference_reductions.npz")
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    for i in range(4):
        eng.upload(i, g["des%d" % i])
        kps = [types.SimpleNamespace(pt=(float(x), float(y))) for x, y in g["pts%d" % i]]
        eng.upload_keypoint_keys(i, matcher.keypoint_keys(kps))
    prm = _capi.Engine.make_params(dedupe=True, cross_check=True)
    table, count = eng.match_pairs([(0, 1), (0, 2)], prm)
    assert table[0, :count[0]].tolist() == g["cross01"].tolist()
    assert table[1, :count[1]].tolist() == g["bidir02_fwd"].tolist()
    prm = _capi.Engine.make_params(dedupe=True, cross_check=False)
    table, count, rtable, rcount = eng.match_pairs([(0, 1)], prm, want_reverse=True)
    assert table[0, :count[0]].tolist() == g["basic01"].tolist()
    assert rtable[0, :rcount[0]].tolist() == g["basic10"].tolist()


def test_match_images_one_call_equals_two_phase():
    """iam_match_images (uploads enqueued wave by wave) == upload everything, then iam_match_pairs."""
    from imageanalysis_b200 import matcher
    des, pts, _ = synth.sift_project(40, 600, seed=21)
    pairs = [(i, j) for i in range(40) for j in range(i + 1, min(40, i + 5))]
    keys = [matcher.keypoint_keys([types.SimpleNamespace(pt=(float(x), float(y))) for x, y in p]) for p in pts]
    prm = _capi.Engine.make_params(dedupe=True, cross_check=True)
    ref = _capi.Engine(_capi.NORM_L2, 128, 0)
    for i, d in enumerate(des):
        ref.upload(i, d.astype(np.float32))
        ref.upload_keypoint_keys(i, keys[i])
    t0, c0 = ref.match_pairs(pairs, prm)
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    for rep in range(2):      # second call re-uploads over resident images (WAR ordering between calls)
        t1, c1 = eng.match_images(list(range(40)), [d.astype(np.float32) for d in des], pairs, prm, keys=keys)
        assert (c0 == c1).all() and c0.max() > 100
        for p in range(len(pairs)):
            assert (t0[p, :c0[p]] == t1[p, :c1[p]]).all()
    assert eng.descriptors_exact(3) == 1 and eng.num_descriptors(39) == 600
    t2, c2 = eng.match_images(list(range(40)), des, pairs, prm)          # uint8 input, no keys
    t3, c3 = ref.match_pairs(pairs, _capi.Engine.make_params(dedupe=False, cross_check=True))
    assert (c2 == c3).all()


@pytest.mark.parametrize("mode", ["0", "1", "2", "1:0.5", "1:0"])
def test_match_images_host_narrowing(monkeypatch, mode):
    """float32 descriptors are narrowed to bytes on the host for transport (IAM_HOST_NARROW: 0 off, 1 adaptive,
    2 always): identical tables in every mode; a non-integer component is never narrowed away -- the call falls back
    to fp16 operands as it does without narrowing."""
    frac = None
    if ":" in mode:          # the planned split: this fraction of the images goes to the workers, the rest crosses as float32
        mode, frac = mode.split(":")
        monkeypatch.setenv("IAM_NARROW_FRACTION", frac)
    monkeypatch.setenv("IAM_HOST_NARROW", mode)
    des, _, _ = synth.sift_project(24, 700, seed=5)
    pairs = [(i, j) for i in range(24) for j in range(i + 1, min(24, i + 5))]
    prm = _capi.Engine.make_params(cross_check=True)
    ref = _capi.Engine(_capi.NORM_L2, 128, 0)
    for i, d in enumerate(des):
        ref.upload(i, d)
    t0, c0 = ref.match_pairs(pairs, prm)
    ref.close()
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    for rep in range(2):
        t1, c1 = eng.match_images(list(range(24)), [d.astype(np.float32) for d in des], pairs, prm)
        assert eng.timing().mma_kind == _capi.KIND_I8
        assert (c0 == c1).all() and c0.max() > 100
        for p in range(len(pairs)):
            assert (t0[p, :c0[p]] == t1[p, :c1[p]]).all()
        if mode == "2" and (os.cpu_count() or 1) >= 2 and not os.environ.get("IAM_HOST_THREADS"):
            assert eng.timing().narrowed_images == 24 and eng.timing().h2d_bytes == 24 * 700 * 128
        if mode == "0" or frac == "0":
            assert eng.timing().narrowed_images == 0 and eng.timing().h2d_bytes == 24 * 700 * 128 * 4
        if frac == "0.5":
            n = eng.timing().narrowed_images
            assert n <= 12 and eng.timing().h2d_bytes == (n + 4 * (24 - n)) * 700 * 128
    eng.close()
    fl = [d.astype(np.float32) for d in des]
    fl[7][33, 5] += 0.5                       # not an integer: bytes cannot carry it
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    table, count = eng.match_images(list(range(24)), fl, pairs, prm)
    assert eng.timing().mma_kind == _capi.KIND_F16
    chk = [p for p, (a, b) in enumerate(pairs) if 7 in (a, b)][:3]
    for p in chk:
        a, b = pairs[p]
        f, _ = oracle.bidirectional(fl[a], fl[b], oracle.NORM_L2, 0.75, 270.0, threads=4)
        assert table[p, :count[p]].tolist() == f
    eng.close()


class FakeImage:
    def __init__(self, name, des, pts, ned):
        self.name = name
        self.des_list = des
        self.kp_list = [types.SimpleNamespace(pt=(float(p[0]), float(p[1]))) for p in pts]
        self.uv_list = [list(map(float, p)) for p in pts]
        self.match_list = {}
        self.matches_clean = True
        self.desc_timestamp = 0.0
        self._ned = list(map(float, ned))
        self.saved = 0

    def get_camera_pose(self, opt=False):
        return self._ned, [0.0, 0.0, 0.0], [1.0, 0.0, 0.0, 0.0]

    def detect_features(self, scale):
        raise AssertionError("descriptors are preloaded")

    def save_matches(self):
        self.saved += 1
        self.matches_clean = True

    def set_aircraft_yaw_error_estimate(self, v):
        pass


def _configure_matcher(detector="SIFT", gms=False):
    """gms=False: the fixtures reference_find_matches / reference_reductions were recorded with an identity
    matchGMS (tests/golden/make_golden.py); the GMS-live fixtures are exercised with gms=True."""
    from imageanalysis_b200 import matcher
    matcher.gms_enabled = gms
    from imageanalysis_b200.propshim import getNode
    det = getNode("/config/detector", True)
    det.setString("detector", detector)
    det.setFloat("scale", 1.0)
    mn = getNode("/config/matcher", True)
    mn.setFloat("match_ratio", 0.75)
    mn.setFloat("min_pairs", 25)
    cam = getNode("/config/camera", True)
    cam.setInt("width_px", 5472)
    cam.setInt("height_px", 3648)
    matcher.configure()
    return matcher


def test_find_matches_equals_reference_driver():
    """The drop-in find_matches() against the match_list the reference's own
    find_matches produced on the same 7-frame project (golden)."""
    g = load_golden("reference_find_matches.npz")
    matcher = _configure_matcher()
    n = int(g["n"])
    imgs = [FakeImage("frame%02d" % i, g["des%d" % i].astype(np.float32), g["pts%d" % i], g["ned%d" % i])
            for i in range(n)]
    proj = types.SimpleNamespace(image_list=imgs, analysis_dir="/tmp")
    K = np.array([[3666.666504, 0, 2736], [0, 3666.666504, 1824], [0, 0, 1]])
    matcher.find_matches(proj, K, strategy="traditional", transform="homography", sort=False, review=False)
    checked = 0
    for im in imgs:
        assert im.saved == 1
        for other, lst in im.match_list.items():
            want = g["match_%s_%s" % (im.name, other)].tolist()
            assert lst == want, (im.name, other)
            checked += 1
    assert checked == 2 * 18
    # resume rule (matcher.py:945-951): a second call skips finished pairs, retries empty ones
    before = {im.name: {k: list(v) for k, v in im.match_list.items()} for im in imgs}
    matcher.find_matches(proj, K, strategy="traditional")
    assert before == {im.name: im.match_list for im in imgs}


def test_find_matches_loads_images_block_wise_and_flushes_them():
    """The 'traditional' strategy brings at most `host_block_images` new images to the host per device call, keeps the
    earlier ones resident on the device, and drops the features it loaded itself once no later pair needs them
    (the reference's descriptor-cache flush, matcher.py:1012-1026).  Results are those of the one-call path."""
    g = load_golden("reference_find_matches.npz")
    matcher = _configure_matcher()
    n = int(g["n"])

    class LazyImage(FakeImage):
        live, peak, loads = 0, 0, 0

        def __init__(self, i):
            super().__init__("frame%02d" % i, None, [], g["ned%d" % i])
            self.i = i
            self.kp_list = None

        def detect_features(self, scale):          # the .feat / .desc cache hit of Image.detect_features (image.py:287-296)
            pts = g["pts%d" % self.i]
            self.des_list = g["des%d" % self.i].astype(np.float32)
            self.kp_list = [types.SimpleNamespace(pt=(float(p[0]), float(p[1]))) for p in pts]
            self.uv_list = [list(map(float, p)) for p in pts]
            LazyImage.loads += 1

    def live_now(imgs):
        return sum(im.des_list is not None for im in imgs)

    imgs = [LazyImage(i) for i in range(n)]
    preloaded = imgs[3]
    preloaded.detect_features(1.0)                  # the caller's own copy must survive
    LazyImage.loads = 0
    proj = types.SimpleNamespace(image_list=imgs, analysis_dir="/tmp")
    K = np.array([[3666.666504, 0, 2736], [0, 3666.666504, 1824], [0, 0, 1]])
    old = matcher.host_block_images
    matcher.host_block_images = 2
    peak = [0]
    real = matcher._capi.Engine.match_images

    def spy(self, ids, arrays, pairs, prm, keys=None, out=None):
        assert len(ids) <= 2 or peak[0] == 0       # the first block holds the first pair's two images
        peak[0] = max(peak[0], live_now(imgs))
        return real(self, ids, arrays, pairs, prm, keys=keys, out=out)

    matcher._capi.Engine.match_images = spy
    try:
        matcher.find_matches(proj, K, strategy="traditional")
    finally:
        matcher._capi.Engine.match_images = real
        matcher.host_block_images = old
    assert LazyImage.loads == n - 1                 # every image loaded exactly once
    assert peak[0] <= 3                             # a block's two images + the caller's preloaded one, never all 7
    assert preloaded.des_list is not None and live_now(imgs) == 1
    checked = 0
    for im in imgs:
        for other, lst in im.match_list.items():
            assert lst == g["match_%s_%s" % (im.name, other)].tolist(), (im.name, other)
            checked += 1
    assert checked == 2 * 18


def test_find_matches_with_gms_equals_reference_driver():
    """find_matches() with the GMS stage live against the reference's own find_matches run with
    cv2.xfeatures2d.matchGMS served by its archive GmsMatcher (tests/golden/make_golden_gms.py)."""
    g = load_golden("reference_gms_pipeline.npz")
    matcher = _configure_matcher(gms=True)
    n = int(g["n"])
    imgs = [FakeImage("frame%02d" % i, g["des%d" % i].astype(np.float32), g["pts%d" % i], g["ned%d" % i])
            for i in range(n)]
    proj = types.SimpleNamespace(image_list=imgs, analysis_dir="/tmp")
    K = np.array([[3666.666504, 0, 2736], [0, 3666.666504, 1824], [0, 0, 1]])
    matcher.find_matches(proj, K, strategy="traditional", transform="homography", sort=False, review=False)
    checked = 0
    for im in imgs:
        for other, lst in im.match_list.items():
            assert lst == g["match_%s_%s" % (im.name, other)].tolist(), (im.name, other)
            checked += 1
    assert checked == 2 * 10
    # the single-pair functions run the same stage through iam_gms_filter
    assert matcher.basic_pair_matches(imgs[0], imgs[1]) == g["basic01"].tolist()
    f, r = matcher.bidirectional_pair_matches(imgs[1], imgs[2])
    assert f == g["bidir12_fwd"].tolist() and r == g["bidir12_rev"].tolist()
    matcher.gms_enabled = False


def test_raw_matches_and_pair_functions():
    g = load_golden("reference_reductions.npz")
    matcher = _configure_matcher()
    i1 = FakeImage("a", g["des0"].astype(np.float32), g["pts0"], [0, 0, 0])
    i2 = FakeImage("b", g["des1"].astype(np.float32), g["pts1"], [15, 0, 0])
    m = matcher.raw_matches(i1, i2, k=3)
    oi, od = oracle.knn(g["des0"], g["des1"], 3, oracle.NORM_L2, threads=4)
    assert len(m) == len(oi) and all(len(r) == 3 for r in m)
    assert [[d.trainIdx for d in r] for r in m] == oi.tolist()
    assert np.float32([[d.distance for d in r] for r in m]).tolist() == od.tolist()
    assert m[5][0].queryIdx == 5
    assert matcher.basic_pair_matches(i1, i2) == g["basic01"].tolist()
    f, r = matcher.bidirectional_pair_matches(i1, i2)
    assert f == g["cross01"].tolist() and r == g["cross10"].tolist()
    empty = FakeImage("c", None, [], [0, 0, 0])
    assert matcher.raw_matches(i1, empty) == []


def test_non_integer_descriptors_tolerance_path():
    """SURF/RootSIFT-like float descriptors go through fp16 operands: distances
    agree with float64 arithmetic to 2e-3 relative (tolerance path)."""
    rng = np.random.default_rng(8)
    q = np.abs(rng.normal(size=(400, 128))).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    t = q[rng.permutation(400)] + rng.normal(0, 0.02, (400, 128)).astype(np.float32)
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    eng.upload(0, q)
    eng.upload(1, t.astype(np.float32))
    assert eng.descriptors_exact(0) == 0
    idx, dist, _, _ = eng.knn_pairs([(0, 1)], 2, 400, reverse=False)
    oi, od = oracle.knn(q, t.astype(np.float32), 2, oracle.NORM_L2)
    assert (idx[0, :, 0] == oi[:, 0]).mean() > 0.99
    assert np.allclose(dist[0, :, 0], od[:, 0], rtol=2e-3, atol=2e-3)


def test_non_integer_descriptors_ratio_and_tables():
    """The tolerance path end to end (north_star: "within a stated tolerance on distance AND ratio"): non-integer float
    descriptors of SIFT magnitude through fp16 operands.  Stated tolerance: distance 2e-3 relative, Lowe ratio d0/d1
    4e-3 absolute on rows whose two neighbours agree (>= 99 % of rows), and the match tables of both reductions agree
    with float64 arithmetic as SETS with IoU >= 0.97 -- every pair that differs sits within 1 % of a threshold."""
    rng = np.random.default_rng(18)
    n = 1500
    base = synth.sift_like(n, seed=77).astype(np.float32)
    q = base + rng.uniform(-0.5, 0.5, base.shape).astype(np.float32)              # non-integer: no exact byte path
    perm = rng.permutation(n)
    t = (base[perm] + rng.normal(0, 2.5, base.shape)).astype(np.float32)
    t[n // 2:] = synth.sift_like(n - n // 2, seed=78).astype(np.float32) + 0.25     # half of the rows have no partner
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    eng.upload(0, q)
    eng.upload(1, t)
    assert eng.descriptors_exact(0) == 0 and eng.descriptors_exact(1) == 0
    idx, dist, ridx, rdist = eng.knn_pairs([(0, 1)], 2, n, reverse=True)
    oi, od = oracle.knn(q, t, 2, oracle.NORM_L2)
    same = (idx[0] == oi).all(axis=1)
    assert same.mean() >= 0.99
    assert np.allclose(dist[0][same], od[same], rtol=2e-3, atol=0)
    r_gpu, r_ref = dist[0][same, 0] / dist[0][same, 1], od[same, 0].astype(np.float64) / od[same, 1]
    assert np.abs(r_gpu - r_ref).max() <= 4e-3
    for mode, name in ((_capi.REDUCE_REF_METRIC, "ref_metric"), (_capi.REDUCE_LOWE, "lowe")):
        prm = _capi.Engine.make_params(reduce_mode=mode, cross_check=True, min_pairs=25, cap=2000)
        table, count = eng.match_pairs([(0, 1)], prm)
        got = {tuple(r) for r in table[0, :count[0]].tolist()}
        f, _ = oracle.bidirectional(q, t, oracle.NORM_L2, 0.75, 270.0, cap=2000, min_pairs=25, threads=4, mode=name, cross_check=True)
        want = {tuple(r) for r in f}
        assert len(want) > 300
        iou = len(got & want) / len(got | want)
        assert iou >= 0.97, (name, iou)
        for qi, ti in got ^ want:         # every disagreement is a row on a threshold
            d0, d1 = float(od[qi, 0]), float(od[qi, 1])
            ratio = d0 / d1
            metric = d0 * ratio
            near = abs(ratio - 0.75) < 0.0075 if name == "lowe" else abs(metric - 270.0 * 0.75) < 2.1
            # ... or its reverse-direction partner is (cross-check removes both)
            if not near:
                ro, rd = oracle.knn(t[ti:ti + 1], q, 2, oracle.NORM_L2)
                e0, e1 = float(rd[0, 0]), float(rd[0, 1])
                near = abs(e0 / e1 - 0.75) < 0.0075 if name == "lowe" else abs(e0 * e0 / e1 - 270.0 * 0.75) < 2.1
            assert near, (name, qi, ti, ratio, metric)


def test_error_paths():
    eng = _capi.Engine(_capi.NORM_L2, 128, 0)
    with pytest.raises(_capi.IamError, match="no descriptors"):
        eng.knn_pairs([(0, 1)], 2, 10)
    eng.upload(0, synth.sift_like(10, 1))
    eng.upload(1, synth.sift_like(10, 2))
    with pytest.raises(_capi.IamError, match="k=4"):
        eng.knn_pairs([(0, 1)], 4, 10)
    with pytest.raises(_capi.IamError):
        eng.upload(2, np.zeros((4, 64), np.uint8))
    eng.release(1)
    assert eng.num_descriptors(1) < 0
    with pytest.raises(_capi.IamError):
        _capi.Engine(7, 128, 0)
