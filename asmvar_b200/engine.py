"""Python host mirror of the reference's AGEaligner interface, on top of the C ABI (include/age_b200.h).

The reference exposes the hot path as `AGEaligner(s1, s2).align(scorer, flag)` + `score()` + `printAlignment()`
(AGEaligner.hh:100-105) and as the serial candidate loop of age_align.cpp:231-293.  `Engine.align()` is that loop
executed as one GPU batch; `Engine.aligner_summary()` is one AGEaligner call.  Everything is computed by the CUDA
library: if libage_b200.so or a CUDA device is missing this module raises, it never falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Sequence

from . import capi


class AgeError(RuntimeError):
    pass


class Engine:
    def __init__(self, devices: Sequence[int] | None = None):
        self.lib = capi.load()
        if devices:
            arr = (C.c_int * len(devices))(*devices)
            self.ctx = self.lib.age_create(arr, len(devices))
        else:
            self.ctx = self.lib.age_create(None, 0)
        if not self.ctx:
            raise AgeError("age_create failed: " + (self.lib.age_last_error(None) or b"").decode())

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.age_destroy(self.ctx)
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def make_jobs(jobs: Iterable[tuple]):
        """jobs: (seq1: bytes, seq2: bytes, flag, match, mismatch, gap_open, gap_extend[, partner]) -> (ctypes array, keepalive, n)"""
        jobs = list(jobs)
        arr = (capi.AgeJob * max(1, len(jobs)))()
        keep = []
        for i, job in enumerate(jobs):
            s1, s2, flag, m, mm, go, ge = job[:7]
            keep.append((s1, s2))
            j = arr[i]
            j.seq1, j.len1, j.seq2, j.len2 = s1, len(s1), s2, len(s2)
            j.flag, j.match, j.mismatch, j.gap_open, j.gap_extend = flag, m, mm, go, ge
            j.partner = job[7] if len(job) > 7 else -1
        return arr, keep, len(jobs)

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise AgeError(f"{what} failed ({rc}): " + (self.lib.age_last_error(self.ctx) or b"").decode())

    def align_raw(self, arr, n: int, out=None):
        """age_align_batch over a prepared job array; returns the ctypes result array (valid until the next call).
        `out` may be a result array from an earlier call, to keep allocation out of a timed loop."""
        res = out if out is not None else (capi.AgeResult * max(1, n))()
        self._check(self.lib.age_align_batch(self.ctx, arr, n, res), "age_align_batch")
        return res

    def align(self, jobs: Iterable[tuple]) -> list[dict]:
        arr, keep, n = self.make_jobs(jobs)
        res = self.align_raw(arr, n)
        return [res[i].summary() for i in range(n)]

    def aligner_summary(self, seq1: bytes, seq2: bytes, flag: int, match: int, mismatch: int, go: int, ge: int) -> dict:
        return self.align([(seq1, seq2, flag, match, mismatch, go, ge)])[0]

    # streaming form: up to two batches in flight (age_submit / age_collect)
    def submit(self, arr, n: int):
        self._check(self.lib.age_submit(self.ctx, arr, n), "age_submit")

    def collect(self, n: int, out=None):
        """Results of the oldest batch in flight (block pointers valid until the second submit after its own)."""
        res = out if out is not None else (capi.AgeResult * max(1, n))()
        self._check(self.lib.age_collect(self.ctx, res, n), "age_collect")
        return res

    def in_flight(self) -> int:
        return self.lib.age_in_flight(self.ctx)

    def align_stream(self, batches: Iterable[Iterable[tuple]]):
        """The candidate loop of age_align.cpp:231-293 over a stream of batches: yields one list of result
        summaries per batch, in order, with the next batch already on the device while one is computed."""
        pending = []
        for jobs in batches:
            arr, keep, n = self.make_jobs(jobs)
            self.submit(arr, n)
            pending.append(n)
            if len(pending) == 2:
                n0 = pending.pop(0)
                res = self.collect(n0)
                yield [res[i].summary() for i in range(n0)]
        while pending:
            n0 = pending.pop(0)
            res = self.collect(n0)
            yield [res[i].summary() for i in range(n0)]

    # split form used by bench.py (inputs resident in HBM when the timed region starts)
    def stage(self, arr, n: int):
        self._check(self.lib.age_stage(self.ctx, arr, n), "age_stage")

    def run_staged(self):
        self._check(self.lib.age_run_staged(self.ctx), "age_run_staged")

    def run_staged_many(self, reps: int) -> float:
        """`reps` runs of the staged batch queued back to back; returns the CUDA-event time of all of them in ms"""
        ms = C.c_double(0.0)
        self._check(self.lib.age_run_staged_many(self.ctx, reps, C.byref(ms)), "age_run_staged_many")
        return ms.value

    def fetch(self, n: int):
        res = (capi.AgeResult * max(1, n))()
        self._check(self.lib.age_fetch(self.ctx, res, n), "age_fetch")
        return res

    def stats(self) -> dict:
        st = capi.AgeStats()
        self._check(self.lib.age_get_stats(self.ctx, C.byref(st)), "age_get_stats")
        return {k: getattr(st, k) for k, _ in capi.AgeStats._fields_}
