"""Parity at the shapes of the BASELINE.json configurations: a handful of full-size candidates per
configuration against the CPU oracle, plus size-independent properties on larger batches where the oracle would take
too long."""
import numpy as np
import pytest

import ageutil as U
from asmvar_b200 import capi, synth
from test_gpu_parity import KEYS, diff_summary

pytestmark = pytest.mark.gpu
SC = (1, -1, -10, -1)


@pytest.fixture(scope="module")
def eng():
    from asmvar_b200 import engine
    with engine.Engine() as e:
        yield e


def both_jobs(cands, flag):
    jobs = []
    for w, q in cands:
        rc = synth.revcom(np.frombuffer(q, dtype=np.uint8)).tobytes()
        k = len(jobs)
        jobs += [(w, q, flag, *SC, k + 1), (w, rc, flag, *SC, k)]
    return jobs


_ORACLE = {}


def oracle(job):
    """the oracle's answer for a job, computed once per module run (a 50 kb x 10 kb job takes it ten seconds)"""
    if job not in _ORACLE:
        _ORACLE[job] = U.oracle_aligner(*job)
    return _ORACLE[job]


def check_both_against_oracle(eng, cands, flag):
    jobs = both_jobs(cands, flag)
    res = eng.align(jobs)
    for i in range(0, len(jobs), 2):
        wa, wb = oracle(jobs[i][:7]), oracle(jobs[i + 1][:7])
        a, b = res[i], res[i + 1]
        assert (a["score"], b["score"]) == (wa["score"], wb["score"])
        win, want = (a, wa) if wa["score"] >= wb["score"] else (b, wb)
        assert win["detail"] == 1 and not diff_summary(win, want), (i, diff_summary(win, want))


def test_c2_indel_5kb(eng):
    check_both_against_oracle(eng, synth.indel_candidates(24, seed=1002), capi.INDEL_FLAG)


def test_c3_tdup_bins(eng):
    cands = []
    for k, (w, f) in enumerate([(1000, 250), (2000, 500), (4000, 500), (8000, 1000)]):
        cands += synth.tdup_candidates(3, seed=1003 + k, window=w, flank=f)
    check_both_against_oracle(eng, cands, capi.TDUPLICATION_FLAG)
    check_both_against_oracle(eng, synth.indel_candidates(4, seed=1003, window=8000, flank=1000), capi.INDEL_FLAG)


@pytest.mark.parametrize("flag", [capi.INVERSION_FLAG, capi.INVL_FLAG, capi.INVR_FLAG])
def test_c4_inversions_20kb(eng, flag):
    """2 kb flanks vs 20 kb windows, with and without -revcom2 (the contig reverse-complemented by the caller)"""
    cands = synth.inversion_candidates(2, seed=1004)
    jobs = []
    for w, q in cands:
        jobs.append((w, q, flag, *SC))
        jobs.append((w, synth.revcom(np.frombuffer(q, dtype=np.uint8)).tobytes(), flag, *SC))
    res = eng.align(jobs)
    for j, r in zip(jobs, res):
        want = U.oracle_aligner(*j)
        assert not diff_summary(r, want), diff_summary(r, want)


@pytest.mark.parametrize("knob", [{}, {"AGE_LONG_NW": "16", "AGE_LONG_CS": "2"}, {"AGE_LONG_NW": "16", "AGE_LONG_CS": "1"}])
def test_c5_long_sv_50kb(eng, knob, monkeypatch):
    """40 strips per pair: by default two pairs get clusters of 4 CTAs (64 warps per pair); forced: clusters of 2, one CTA"""
    for k, v in knob.items():
        monkeypatch.setenv(k, v)
    check_both_against_oracle(eng, synth.long_sv_candidates(2, seed=1005), capi.INDEL_FLAG)


def test_full_batch_properties(eng):
    """512 C2 candidates (the oracle would need minutes): results do not depend on the batch composition, the -both
    winner carries breakpoints, fragments lie inside the sequences and the score bounds hold."""
    cands = synth.indel_candidates(512, seed=77)
    jobs = both_jobs(cands, capi.INDEL_FLAG)
    res = eng.align(jobs)
    sub = list(range(0, 64, 2))
    res2 = eng.align([jobs[i][:7] + (-1,) for i in sub] + [jobs[i + 1][:7] + (-1,) for i in sub])   # no partners, other pairing
    for n, i in enumerate(sub):
        assert res2[n]["score"] == res[i]["score"] and res2[len(sub) + n]["score"] == res[i + 1]["score"]
        win = res[i] if res[i]["detail"] else res[i + 1]
        alone = res2[n] if res[i]["detail"] else res2[len(sub) + n]
        assert not diff_summary(alone, {k: win[k] for k in KEYS if k in win})
    for i in range(0, len(jobs), 2):
        a, b = res[i], res[i + 1]
        assert a["status"] == 0 and b["status"] == 0 and a["detail"] + b["detail"] == 1
        win, job = (a, jobs[i]) if a["detail"] else (b, jobs[i + 1])
        assert win["score"] >= max(a["score"], b["score"]) and 0 < win["score"] <= min(len(job[0]), len(job[1]))
        assert 1 <= len(win["bp"]) <= 100
        for s1, s2, n in win["left"] + win["right"]:
            assert 1 <= s1 and s1 + n - 1 <= len(job[0]) and 1 <= s2 and s2 + n - 1 <= len(job[1]) and n >= 1
        assert sum(n for _, _, n in win["left"] + win["right"]) >= win["score"]          # match = 1: aligned >= score
