"""Host-side planning on the CPU (age_plan, no CUDA device needed): the longest-processing-time-first deal of pairs to the
devices of one context (csrc/age_engine.cu: deal_pairs, used by make_subbatches), the packed / wide classification of
every job, and -- with two gloo ranks -- that ranks which plan independently agree and that a host-side gather restores
input order (the only "collective" of the path)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

from asmvar_b200 import capi
from asmvar_b200.engine import Engine


def plan(jobs, n_dev):
    arr, keep, n = Engine.make_jobs(jobs)
    dev = (C.c_int32 * max(1, n))()
    ar = (C.c_int32 * max(1, n))()
    assert capi.load().age_plan(arr, n, n_dev, dev, ar) == 0
    return list(dev[:n]), list(ar[:n])


def rand_jobs(seed, n, both=True):
    rng = np.random.default_rng(seed)
    jobs = []
    for i in range(n):
        l1, l2 = int(rng.integers(200, 8000)), int(rng.integers(100, 2000))
        w, q = bytes(rng.choice(list(b"ACGT"), l1).astype(np.uint8)), bytes(rng.choice(list(b"ACGT"), l2).astype(np.uint8))
        k = len(jobs)
        if both:
            jobs += [(w, q, capi.INDEL_FLAG, 1, -1, -10, -1, k + 1), (w, q[::-1], capi.INDEL_FLAG, 1, -1, -10, -1, k)]
        else:
            jobs.append((w, q, capi.TDUPLICATION_FLAG, 1, -1, -10, -1, -1))
    return jobs


@pytest.mark.parametrize("n_dev", [1, 2, 4, 8])
def test_deal_is_a_balanced_partition(n_dev):
    jobs = rand_jobs(3, 300)
    dev, ar = plan(jobs, n_dev)
    assert all(0 <= d < n_dev for d in dev) and set(ar) == {16}
    for k in range(0, len(jobs), 2):                      # -both partners share a warp, hence a device
        assert dev[k] == dev[k + 1]
    cost = [len(j[0]) * len(j[1]) for j in jobs]
    load = [sum(c for c, d in zip(cost, dev) if d == r) for r in range(n_dev)]
    assert max(load) - min(load) <= 2 * max(cost)         # LPT: within one pair (two jobs) of each other
    assert plan(jobs, n_dev) == (dev, ar)                 # deterministic


def test_unpaired_jobs_are_paired_by_shape_and_dealt():
    jobs = rand_jobs(4, 64, both=False) * 2                # every shape twice: identical shapes share a warp
    dev, _ = plan(jobs, 4)
    assert dev[:64] == dev[64:]
    assert len(set(dev)) == 4


def test_arithmetic_classification_and_status():
    A = b"ACGT" * 50
    L = b"ACGT" * 4200                                     # 16.8 kb
    jobs = [
        (A, A, capi.INDEL_FLAG, 1, -1, -10, -1),           # packed
        (A, A, capi.INDEL_FLAG, 63, -63, -40, -2),         # packed: 2*63*200 + ... < 32767
        (A, A, capi.INDEL_FLAG, 64, -1, -10, -1),          # |match| > 63: wide
        (A, A, capi.INDEL_FLAG, 1, -1, 0, -1),             # gap open >= 0: wide
        (A, A, capi.INDEL_FLAG, 1, -1, -10, 1),            # gap extend >= 0: wide
        (L, L, capi.INDEL_FLAG, 1, -1, -10, -1),           # match 1 x 16.8 kb > 16 k: wide
        (L, L[:8000], capi.INDEL_FLAG, 2, -2, -10, -1),    # match 2 x 8 kb: scores to 16 000 still fit 2*score + tag <= 32767
        (L, L[:8200], capi.INDEL_FLAG, 2, -2, -10, -1),    # just beyond
        (L, L[:5000], capi.INDEL_FLAG, 1, 4, -10, -1),     # a mismatch that scores more than a match bounds the range (4 x 5000)
        (A, A, capi.INDEL_FLAG, 20000, -1, -10, -1),       # 2*match does not fit the wide score tables
        (A, b"", capi.INDEL_FLAG, 1, -1, -10, -1),         # AGEaligner.cpp:68
        (A, A, 0, 1, -1, -10, -1),                         # AGEaligner.cpp:69
        (A, A, capi.INDEL_FLAG | capi.INVL_FLAG, 1, -1, -10, -1),
        (A, A, capi.INVERSION_FLAG | capi.INVL_FLAG, 1, -1, -10, -1),   # INVERSION wins (AGEaligner.cpp:72-77)
    ]
    dev, ar = plan(jobs, 2)
    assert ar == [16, 16, 32, 32, 32, 32, 16, 32, 32, capi.AGE_ERR_RANGE, capi.AGE_NOT_ALIGNED, capi.AGE_NOT_ALIGNED, capi.AGE_ERR_ARG, 16]
    assert [d >= 0 for d in dev] == [a >= 16 for a in ar]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    jobs = rand_jobs(9, 40)
    dev, _ = plan(jobs, world)                             # every rank plans by itself: no communication
    mine = [i for i, d in enumerate(dev) if d == rank]
    records = [(i, f"record-{i}") for i in mine]           # stands for the results of this rank's jobs
    parts = [None] * world
    dist.all_gather_object(parts, records)                 # host-side gather
    out = [None] * len(jobs)
    for part in parts:
        for i, rec in part:
            out[i] = rec
    q.put((rank, mine, out))
    dist.destroy_process_group()


def test_two_ranks_plan_alike_and_gather_in_input_order():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = {r: mine for r, mine, _ in res}
    assert sorted(shards[0] + shards[1]) == list(range(80)) and not set(shards[0]) & set(shards[1])
    for _, _, out in res:
        assert out == [f"record-{i}" for i in range(80)]
