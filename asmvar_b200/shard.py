"""Host-side sharding of candidates across ranks (one process per GPU).

Candidates are independent -- the reference exposes this as `--selectId` per-target process sharding
(age_align.cpp:233) -- so there is no data-path collective: every rank aligns its own shard and only the
per-candidate results (or, in bench.py, the per-rank counters) are gathered on the host.
The deal is longest-processing-time-first on the DP area len1*len2, the same policy the C engine uses between the
devices of one process (csrc/age_engine.cu: make_subbatches).
"""
from __future__ import annotations

from typing import Sequence


def lpt_partition(costs: Sequence[float], world: int) -> list[list[int]]:
    """Indices of `costs` dealt to `world` shards, largest first onto the least-loaded shard; each shard keeps
    input order.  Deterministic: every rank computes the same partition without communication."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    shards: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        k = min(range(world), key=lambda r: (load[r], r))
        shards[k].append(i)
        load[k] += costs[i]
    for s in shards:
        s.sort()
    return shards


def gather_in_order(local_items: list, local_index: list[int], total: int, world: int):
    """all_gather_object of (index, item) pairs -> list of `total` items in input order (host-side gather)."""
    import torch.distributed as dist
    if world == 1:
        out = [None] * total
        for i, it in zip(local_index, local_items):
            out[i] = it
        return out
    parts = [None] * world
    dist.all_gather_object(parts, list(zip(local_index, local_items)))
    out = [None] * total
    for part in parts:
        for i, it in part:
            out[i] = it
    return out
