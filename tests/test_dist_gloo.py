"""World-size-2 test of the pair sharding + table all-gather on the gloo
backend (CPU), exercising the same code the NCCL path runs."""
import os
import socket
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, cap, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from imageanalysis_b200 import dist
    r, w, _ = dist.init("gloo")
    assert (r, w) == (rank, world)
    pairs = np.stack([np.arange(n_total), np.arange(n_total) + 1], 1).astype(np.int32)
    local, b, e = dist.shard_pairs(pairs, rank, world)
    # fake per-pair tables: pair p has (p % cap) entries [p, e]
    table = torch.zeros((e - b, cap, 2), dtype=torch.int32)
    count = torch.zeros((e - b,), dtype=torch.int32)
    for li, p in enumerate(range(b, e)):
        c = p % cap
        count[li] = c
        for k in range(c):
            table[li, k, 0] = p
            table[li, k, 1] = k
    t, c = dist.allgather_tables(table, count, n_total, rank, world)
    torch.save((t, c), os.path.join(out_dir, "r%d.pt" % rank))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_allgather_tables_world2(tmp_path):
    world, n_total, cap = 2, 11, 5
    mp.spawn(_worker, args=(world, _free_port(), n_total, cap, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, "r%d.pt" % r)) for r in range(world)]
    for t, c in res:
        assert t.shape == (n_total, cap, 2) and c.tolist() == [p % cap for p in range(n_total)]
        for p in range(n_total):
            for k in range(p % cap):
                assert t[p, k].tolist() == [p, k]
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])   # byte-identical on all ranks


def test_allgather_tables_world2_equal_shards(tmp_path):
    # pair count divisible by the world size: shards are gathered in place (no staging copies)
    world, n_total, cap = 2, 12, 5
    mp.spawn(_worker, args=(world, _free_port(), n_total, cap, str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(tmp_path, "r%d.pt" % r)) for r in range(world)]
    for t, c in res:
        assert t.shape == (n_total, cap, 2) and c.tolist() == [p % cap for p in range(n_total)]
        for p in range(n_total):
            for k in range(p % cap):
                assert t[p, k].tolist() == [p, k]
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])


def _worker_packed(rank, world, port, n_total, cap, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from imageanalysis_b200 import dist
    dist.init("gloo")
    pairs = np.stack([np.arange(n_total), np.arange(n_total) + 1], 1).astype(np.int32)
    local, b, e = dist.shard_pairs(pairs, rank, world)
    rows, count = [], []
    for p in range(b, e):          # pair p has (p * 7) % cap rows [p, k]: ragged, some empty
        c = (p * 7) % cap
        count.append(c)
        rows += [[p, k] for k in range(c)]
    rows = torch.tensor(rows, dtype=torch.int32).reshape(-1, 2)
    count = torch.tensor(count, dtype=torch.int32)
    r_all, c_all = dist.allgather_packed(rows, count, n_total, rank, world)
    torch.save((r_all, c_all), os.path.join(out_dir, "p%d.pt" % rank))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_allgather_packed_world2(tmp_path):
    """The compact gather (counts, then one padded payload): CSR of every pair's rows, in work-list order, on all ranks."""
    for n_total in (11, 12, 3):
        world, cap = 2, 5
        mp.spawn(_worker_packed, args=(world, _free_port(), n_total, cap, str(tmp_path)), nprocs=world, join=True)
        res = [torch.load(os.path.join(tmp_path, "p%d.pt" % r)) for r in range(world)]
        want_c = [(p * 7) % cap for p in range(n_total)]
        want_r = [[p, k] for p in range(n_total) for k in range((p * 7) % cap)]
        for r_all, c_all in res:
            assert c_all.tolist() == want_c
            assert r_all.reshape(-1, 2).tolist() == want_r


def test_allgather_single_rank_is_identity():
    sys.path.insert(0, ROOT)
    from imageanalysis_b200 import dist
    t = torch.arange(24, dtype=torch.int32).reshape(3, 4, 2)
    c = torch.tensor([1, 2, 3], dtype=torch.int32)
    t2, c2 = dist.allgather_tables(t, c, 3, 0, 1)
    assert t2 is t and c2 is c


def test_bates_worklist_matches_survey_count():
    """BASELINE configs[3]: 38 x 74 serpentine grid, the reference's camera-distance window (matcher.py:858-903):
    32 neighbours per interior frame, ~4.3e4 pairs; shards tile the list without gaps."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from imageanalysis_b200 import pairs as wl
    p = bench.bates_pairs(2812)
    assert len(p) == 42694 and (p[:, 0] < p[:, 1]).all()
    deg = np.bincount(p.ravel(), minlength=2812)
    assert deg.max() == 32 and np.median(deg) == 32
    assert (np.lexsort((p[:, 1], p[:, 0])) == np.arange(len(p))).all()      # (i, j)-sorted: contiguous shards share image i
    cuts = [wl.shard(len(p), r, 8) for r in range(8)]
    assert cuts[0][0] == 0 and cuts[-1][1] == len(p) and all(cuts[r][1] == cuts[r + 1][0] for r in range(7))
