"""Multi-GPU sharding of a batch: one process per GPU, strings split by contiguous
ranges, the transducer replicated on every GPU, NO collective on the data path
(SURVEY §8e).  The only communication is the optional gather of per-rank results
for callers that want the whole batch on one rank (control plane, gloo or nccl)."""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, world: int):
    """Contiguous, balanced ranges: rank r owns [b[r], b[r+1])."""
    base, rem = divmod(n, world)
    b = [0]
    for r in range(world):
        b.append(b[-1] + base + (1 if r < rem else 0))
    return b


def shard_by_cost(lengths: np.ndarray, world: int, power: float = 2.0):
    """Contiguous ranges balanced by an estimated cost ~ len**power (wide frontiers grow
    quadratically with the input length, reference bench/optimize-bench.zig:265-266)."""
    cost = np.power(np.asarray(lengths, np.float64) + 1.0, power)
    csum = np.concatenate([[0.0], np.cumsum(cost)])
    targets = csum[-1] * np.arange(1, world) / world
    cuts = np.searchsorted(csum, targets, side="left")
    b = [0] + [int(c) for c in cuts] + [len(lengths)]
    for i in range(1, len(b)):
        b[i] = max(b[i], b[i - 1])
    return b


def local_slice(data: np.ndarray, offsets: np.ndarray, lo: int, hi: int):
    """Sub-batch [lo, hi) with offsets rebased to 0."""
    off = offsets[lo:hi + 1].astype(np.uint64)
    return data[int(off[0]):int(off[-1])], off - off[0]


def search_sharded(search_fn, data: np.ndarray, offsets: np.ndarray, rank: int, world: int, gather=None, by_cost=False):
    """Run `search_fn(data, offsets) -> list of per-string results` on this rank's shard.
    With `gather` (e.g. torch.distributed.all_gather_object) returns the whole batch in input
    order on every rank; otherwise only this rank's results and its range."""
    n = len(offsets) - 1
    b = shard_by_cost(np.diff(offsets.astype(np.int64)), world) if by_cost else shard_bounds(n, world)
    lo, hi = b[rank], b[rank + 1]
    d, o = local_slice(data, offsets, lo, hi)
    mine = search_fn(d, o)
    assert len(mine) == hi - lo
    if gather is None:
        return mine, (lo, hi)
    parts = [None] * world
    gather(parts, mine)
    out = []
    for p in parts:
        out.extend(p)
    return out, (lo, hi)
