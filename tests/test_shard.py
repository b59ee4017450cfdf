"""Multi-process host logic on CPU (gloo, world_size 2): the LPT deal is a partition, every rank derives it without
communication, and the host-side gather restores input order."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from asmvar_b200 import shard


def test_lpt_partition_is_balanced_partition():
    costs = [5000 * n for n in (1000, 2000, 1500, 1000, 1100, 1999, 1000, 1000, 1700)]
    for world in (1, 2, 4, 8):
        sh = shard.lpt_partition(costs, world)
        assert sorted(i for s in sh for i in s) == list(range(len(costs)))
        loads = [sum(costs[i] for i in s) for s in sh]
        assert max(loads) - min(loads) <= max(costs)
        assert all(s == sorted(s) for s in sh)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = [(i * 37) % 11 + 1 for i in range(23)]
    mine = shard.lpt_partition(costs, world)[rank]
    items = [f"record-{i}" for i in mine]                  # stands for the per-candidate .age records of this rank
    out = shard.gather_in_order(items, mine, len(costs), world)
    q.put((rank, mine, out))
    dist.destroy_process_group()


def test_two_ranks_gather_in_input_order():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = {r: mine for r, mine, _ in res}
    assert sorted(shards[0] + shards[1]) == list(range(23)) and not set(shards[0]) & set(shards[1])
    for _, _, out in res:
        assert out == [f"record-{i}" for i in range(23)]
