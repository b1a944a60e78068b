"""Multi-GPU plumbing: one process per GPU (torchrun), pair list sharded in
contiguous blocks, ONE all-gather of the per-pair match tables over
NCCL/NVLink at the end.  The reference is single
process (matcher.py:928 walks the work list serially); pairs are independent,
so no other communication exists on this path.

Works with backend "nccl" (GPU tensors) and "gloo" (CPU tensors; used by the
world_size-2 tests that run without a GPU).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as td

from . import pairs as _pairs


def env_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from the torchrun environment (no-op for a
    single process).  Returns (rank, world, local_rank)."""
    rank, world, local = env_world()
    if world > 1 and not td.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
        td.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_pairs(pair_array: np.ndarray, rank: int, world: int) -> Tuple[np.ndarray, int, int]:
    b, e = _pairs.shard(len(pair_array), rank, world)
    return pair_array[b:e], b, e


class DevicePtr:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr: int, shape, typestr: str = "<i4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def as_tensor(ptr: int, shape, device: int) -> torch.Tensor:
    return torch.as_tensor(DevicePtr(ptr, shape), device=torch.device("cuda", device))


def allgather_tables(table: torch.Tensor, count: torch.Tensor, n_total: int, rank: int, world: int, out=None):
    """table [P_local, cap, 2] int32, count [P_local] int32 on every rank (the
    rank's contiguous shard) -> (table [n_total, cap, 2], count [n_total]) on
    every rank; `out` = optional preallocated pair of result tensors.

    Equal shards (the benchmark's weak-scaling layout, and any pair count
    divisible by the world size): the shards are gathered straight into their
    final place -- one all-gather of the tables and one of the 4-byte counts,
    no staging copies.  Ragged shards: counts ride in an extra table row so a
    single all-gather (padded to the largest shard) moves everything."""
    if world == 1:
        return table, count
    cap = table.shape[1]
    sizes = [_pairs.shard(n_total, r, world) for r in range(world)]
    p_max = max(e - b for b, e in sizes)
    if out is not None:
        out_t, out_c = out
    else:
        out_t = torch.empty((n_total, cap, 2), dtype=torch.int32, device=table.device)
        out_c = torch.empty((n_total,), dtype=torch.int32, device=table.device)
    if all(e - b == p_max for b, e in sizes):
        td.all_gather_into_tensor(out_t, table.contiguous())   # THE collective of this path
        td.all_gather_into_tensor(out_c, count.contiguous())
        return out_t, out_c
    send = torch.zeros((p_max, cap + 1, 2), dtype=torch.int32, device=table.device)
    p_loc = table.shape[0]
    send[:p_loc, :cap] = table
    send[:p_loc, cap, 0] = count
    recv = torch.empty((world * p_max, cap + 1, 2), dtype=torch.int32, device=table.device)
    td.all_gather_into_tensor(recv, send)      # THE collective of this path
    recv = recv.view(world, p_max, cap + 1, 2)
    for r, (b, e) in enumerate(sizes):
        out_t[b:e] = recv[r, :e - b, :cap]
        out_c[b:e] = recv[r, :e - b, cap, 0]
    return out_t, out_c


def allgather_packed(rows: torch.Tensor, count: torch.Tensor, n_total: int, rank: int, world: int):
    """The gather in compact form ("all-gather counts first, then a variable-size payload
    padded to the per-rank maximum"): rows [T_local, 2] int32 = the valid rows of this rank's pairs back to back
    (iam_pack_tables_device), count [P_local] int32.  Returns (rows_all [T_total, 2], count_all [n_total]) on every
    rank, pairs in work-list order -- the CSR form of `match_list` (offsets = exclusive cumsum of count_all).
    Two collectives: the 4-byte counts, then ONE all-gather of the row payload.  Mean fill of the padded tables is
    a few per cent to ~40 %, so this moves a fraction of the bytes of allgather_tables()."""
    if world == 1:
        return rows, count
    sizes = [_pairs.shard(n_total, r, world) for r in range(world)]
    p_max = max(e - b for b, e in sizes)
    dev = count.device
    send_c = count if count.shape[0] == p_max else torch.cat(
        [count, torch.zeros(p_max - count.shape[0], dtype=count.dtype, device=dev)])
    recv_c = torch.empty((world * p_max,), dtype=torch.int32, device=dev)
    td.all_gather_into_tensor(recv_c, send_c.contiguous())
    recv_c = recv_c.view(world, p_max)
    totals = recv_c.sum(dim=1).cpu().tolist()          # the one host round trip: payload sizes
    t_max = max(1, int(max(totals)))
    send_r = torch.empty((t_max, 2), dtype=torch.int32, device=dev)
    send_r[:rows.shape[0]] = rows
    recv_r = torch.empty((world * t_max, 2), dtype=torch.int32, device=dev)
    td.all_gather_into_tensor(recv_r, send_r)           # THE collective of this path (payload)
    recv_r = recv_r.view(world, t_max, 2)
    rows_all = torch.cat([recv_r[r, :int(totals[r])] for r in range(world)], dim=0)
    count_all = torch.cat([recv_c[r, :e - b] for r, (b, e) in enumerate(sizes)], dim=0)
    return rows_all, count_all
