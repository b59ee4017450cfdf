"""age_submit / age_collect: two batches in flight give the results of age_align_batch, in order.

Covers what the streaming form adds on the host side: slot reuse, the shared scratch arena growing while a batch
is in flight, the block-pool overflow re-run of a batch whose scratch a later batch has already reused, and the
argument errors of the protocol."""
import numpy as np
import pytest

from asmvar_b200 import capi, engine
from test_gpu_parity import random_jobs, diff_summary, check_against_oracle  # noqa: F401

pytestmark = pytest.mark.gpu


def batches():
    # sizes chosen so that the scratch arena has to grow while the previous batch is still in flight
    return [random_jobs(31, 24, (80, 300), (20, 60), ["ACGT", "ACGTN"]),
            random_jobs(32, 40, (900, 2200), (200, 420), ["ACGT"]),
            random_jobs(33, 8, (60, 200), (15, 40), ["ACGTacgtNnU"]),
            random_jobs(34, 32, (300, 1500), (60, 300), ["ACGT", "ACGTN"]),
            random_jobs(35, 16, (80, 300), (20, 60), ["ACGT"])]


def test_stream_equals_batch():
    bs = batches()
    with engine.Engine([0]) as e1:
        want = [e1.align(b) for b in bs]
    with engine.Engine([0]) as e2:
        got = list(e2.align_stream(bs))
        assert e2.in_flight() == 0
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert len(g) == len(w)
        for x, y in zip(g, w):
            assert not diff_summary(x, y)


def test_stream_against_oracle_interleaved():
    """submit A, submit B, collect A, submit C, collect B, collect C -- each checked against the CPU oracle"""
    import ageutil as U
    bs = batches()[:3]
    with engine.Engine([0]) as e:
        prepared = [e.make_jobs(b) for b in bs]
        e.submit(prepared[0][0], prepared[0][2])
        e.submit(prepared[1][0], prepared[1][2])
        assert e.in_flight() == 2
        r0 = [r.summary() for r in e.collect(prepared[0][2])[:prepared[0][2]]]
        e.submit(prepared[2][0], prepared[2][2])
        r1 = [r.summary() for r in e.collect(prepared[1][2])[:prepared[1][2]]]
        r2 = [r.summary() for r in e.collect(prepared[2][2])[:prepared[2][2]]]
    for jobs, res in zip(bs, (r0, r1, r2)):
        for j, r in zip(jobs, res):
            assert not diff_summary(r, U.oracle_aligner(*j))


def test_block_pool_overflow_in_flight(monkeypatch):
    """AGE_POOL_CAP=8: every batch overflows its block pool and is run again at collect time, after the next batch
    has already been launched into the same scratch arena"""
    bs = batches()[:4]
    with engine.Engine([0]) as e1:
        want = [e1.align(b) for b in bs]
    monkeypatch.setenv("AGE_POOL_CAP", "8")
    with engine.Engine([0]) as e2:
        got = list(e2.align_stream(bs))
    for g, w in zip(got, want):
        for x, y in zip(g, w):
            assert not diff_summary(x, y)


def test_protocol_errors():
    jobs = random_jobs(36, 4, (60, 120), (15, 30), ["ACGT"])
    with engine.Engine([0]) as e:
        arr, keep, n = e.make_jobs(jobs)
        with pytest.raises(engine.AgeError):
            e.collect(n)                                   # nothing in flight
        e.submit(arr, n)
        e.submit(arr, n)
        with pytest.raises(engine.AgeError):
            e.submit(arr, n)                               # a third batch needs a collect first
        with pytest.raises(engine.AgeError):
            e.align_raw(arr, n)                            # the synchronous call does not mix with batches in flight
        with pytest.raises(engine.AgeError):
            e.collect(n + 1)                               # wrong size
        a = e.collect(n)
        sa = [a[i].summary() for i in range(n)]
        b = e.collect(n)
        sb = [b[i].summary() for i in range(n)]
        assert e.in_flight() == 0
        for x, y in zip(sa, sb):
            assert not diff_summary(x, y)
        c = e.align(jobs)
        for x, y in zip(sa, c):
            assert not diff_summary(x, y)


def test_stream_two_devices():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    bs = batches()
    with engine.Engine([0]) as e1:
        want = [e1.align(b) for b in bs]
    with engine.Engine([0, 1]) as e2:
        got = list(e2.align_stream(bs))
    for g, w in zip(got, want):
        for x, y in zip(g, w):
            assert not diff_summary(x, y)


def test_batches_larger_than_the_memory_budget(monkeypatch):
    """AGE_MEM_BUDGET_MB=4: every batch is cut into several sub-batches that run one after the other (the path a batch
    larger than the device memory takes), alone and with a normal batch in flight before it"""
    bs = batches()
    with engine.Engine([0]) as e1:
        want = [e1.align(b) for b in bs]
    with engine.Engine([0]) as e2:
        first = e2.make_jobs(bs[0])
        e2.submit(first[0], first[2])                      # in flight, one piece
        monkeypatch.setenv("AGE_MEM_BUDGET_MB", "4")
        second = e2.make_jobs(bs[1])
        e2.submit(second[0], second[2])                    # several pieces: runs synchronously behind the first
        r0 = [r.summary() for r in e2.collect(first[2])[:first[2]]]
        r1 = [r.summary() for r in e2.collect(second[2])[:second[2]]]
        got = [r0, r1] + [e2.align(b) for b in bs[2:]]
        with pytest.raises(engine.AgeError):
            e2.stage(second[0], second[2])                 # the split form needs the batch in one piece
    for g, w in zip(got, want):
        assert len(g) == len(w)
        for x, y in zip(g, w):
            assert not diff_summary(x, y)
