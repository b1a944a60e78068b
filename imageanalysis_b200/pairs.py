"""Pair work-list generation for find_matches (reference
scripts/lib/matcher.py:858-903), vectorised.

The reference walks an O(n^2) Python double loop calling get_camera_pose()
3.95 M times at n = 2812; here the same list — same pairs, same order, same
discretised distance key — comes out of numpy.

  sequential : the live branch, `abs(i-j) <= 4` (matcher.py:899; 4n-10 pairs)
  geotag     : the documented camera-distance window `min_dist <= d <= max_dist`
               (matcher.py:896, switched off by `if False` in the snapshot,
               SURVEY.md D5); default max_dist = 4 x max(median, mean) adjacent
               interval (matcher.py:858-883)
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np


def interval_stats(neds: np.ndarray):
    """matcher.py:858-874 -> (median_int, default_max_dist)."""
    d = np.linalg.norm(neds[1:] - neds[:-1], axis=1) if len(neds) > 1 else np.zeros(0)
    if d.size == 0:
        return 1, 4
    median = float(np.median(d))
    average = float(np.average(d))
    if median < average:
        median = average
    median_int = int(round(median))
    if median_int == 0:
        median_int = 1
    return median_int, median_int * 4


def worklist(neds: Sequence[Sequence[float]], mode: str = "sequential", min_dist: float = 0.0,
             max_dist: Optional[float] = None, seq_k: int = 4) -> List[list]:
    """Returns [[ddist, i, j], ...] with i < j, in the reference's generation
    order (i ascending, then j ascending)."""
    P = np.asarray(neds, np.float64).reshape(-1, 3)
    n = P.shape[0]
    median_int, default_max = interval_stats(P)
    if max_dist is None:
        max_dist = default_max
    interval = median_int * 1.3
    out: List[list] = []
    if mode == "geotag":
        step = 512
        for i0 in range(0, n, step):
            blk = P[i0:i0 + step]
            D = np.sqrt(((blk[:, None, :] - P[None, :, :]) ** 2).sum(-1))
            ii, jj = np.nonzero((D >= min_dist) & (D <= max_dist))
            keep = jj > (ii + i0)
            ii, jj = ii[keep], jj[keep]
            dd = np.rint(D[ii, jj] / interval) * interval
            out.extend([float(a), int(b + i0), int(c)] for a, b, c in zip(dd, ii, jj))
    elif mode == "sequential":
        for i in range(n):
            j = np.arange(i + 1, min(n, i + seq_k + 1))
            if j.size == 0:
                continue
            d = np.sqrt(((P[j] - P[i]) ** 2).sum(-1))
            dd = np.rint(d / interval) * interval
            out.extend([float(a), i, int(b)] for a, b in zip(dd, j))
    else:
        raise ValueError("unknown pair filter '%s' (sequential | geotag)" % mode)
    return out


def pair_array(work: List[list]) -> np.ndarray:
    return np.asarray([[w[1], w[2]] for w in work], np.int32).reshape(-1, 2)


def shard(n_pairs: int, rank: int, world: int):
    """Contiguous block of the (i, j)-sorted pair list for one rank, so that
    consecutive pairs share image i and its descriptor block stays in L2
    (SURVEY.md section 8e).  Returns (begin, end)."""
    base, rem = divmod(n_pairs, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)
