"""Read sharding across the GPUs of one box.

Reads are independent units; the index is replicated; ranks get contiguous, length-balanced ranges of the read
list so that results concatenate back in input order without any data-path collective."""
from __future__ import annotations

import numpy as np


def balanced_ranges(lengths, n_ranks: int):
    """Split reads into n_ranks contiguous ranges balanced by total bases. Returns [(begin, end)] * n_ranks."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n = lengths.size
    if n_ranks <= 0:
        raise ValueError("n_ranks must be positive")
    cum = np.concatenate(([0], np.cumsum(lengths)))
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, n_ranks):
        target = total * r / n_ranks
        b = int(np.searchsorted(cum, target, side="left"))
        b = min(max(b, bounds[-1]), n)
        bounds.append(b)
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(n_ranks)]


def shard(buf: np.ndarray, offsets: np.ndarray, rank: int, n_ranks: int):
    """Rank's slice of a packed read set (bytes, uint64 offsets[n+1]) -> (bytes, offsets, (begin, end))."""
    lengths = np.diff(offsets.astype(np.int64))
    b, e = balanced_ranges(lengths, n_ranks)[rank]
    lo, hi = int(offsets[b]), int(offsets[e])
    return buf[lo:hi], (offsets[b:e + 1] - offsets[b]).astype(np.uint64), (b, e)
