"""The 32-bit ("wide") kernels: every job whose scores do not fit the packed 16-bit halves -- long contigs at match > 1,
gap costs >= 0, |match| > 63 -- and the reference's unsigned-short storage beyond 65535 (AGEaligner.hh:86: scores wrap
modulo 65536 when stored, AGEaligner.cpp:479,548: so does the split sum; the equality tests of findBPs do not).
Compared bit-exactly against the CPU oracle, which keeps the same u16 matrices."""
import numpy as np
import pytest

import ageutil as U
from asmvar_b200 import capi
from test_gpu_parity import CASES, FLAGS, case_jobs, check_against_oracle, mutate, rand_seq, random_jobs, sv_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from asmvar_b200 import engine
    with engine.Engine() as e:
        yield e


def golden_batch_matches(eng):
    jobs, owner = [], []
    for ci, c in enumerate(CASES):
        for j in case_jobs(c, len(jobs), True):
            jobs.append(j); owner.append(ci)
    res = eng.align(jobs)
    by_case = {}
    for r, ci in zip(res, owner):
        by_case.setdefault(ci, []).append(r)
    bad = []
    for ci, c in enumerate(CASES):
        it = iter(by_case[ci])
        got = U.candidate_record(c, lambda *a: next(it))
        if got.rstrip("\n") != c["expected"].rstrip("\n"):
            bad.append(c["id"])
    return bad


@pytest.mark.parametrize("knob", [{}, {"AGE_LONG_NW": "1"}, {"AGE_LONG_NW": "16"}, {"AGE_QUEUE_CAP": "4"}, {"AGE_QUEUE_CAP": "4", "AGE_LONG_NW": "1"}, {"AGE_TRACE_POOL_GRANULES": "400"}, {"AGE_LONG_NW": "8", "AGE_LONG_CS": "2"}])
def test_golden_records_forced_wide(eng, knob, monkeypatch):
    """all 247 reference records through the 32-bit kernels"""
    monkeypatch.setenv("AGE_FORCE_WIDE", "1")
    for k, v in knob.items():
        monkeypatch.setenv(k, v)
    bad = golden_batch_matches(eng)
    assert not bad, f"{knob}: {len(bad)} records differ from the reference: {bad[:20]}"


def test_random_forced_wide(eng, monkeypatch):
    monkeypatch.setenv("AGE_FORCE_WIDE", "1")
    check_against_oracle(eng, random_jobs(31, 120, (60, 400), (15, 80), ["ACGT", "ACGT", "ACGTN", "ACGTacgtNnU"]))
    check_against_oracle(eng, random_jobs(32, 24, (900, 2200), (200, 420), ["ACGT", "ACGTN"]))


def test_scorings_outside_the_packed_range(eng):
    """gap costs >= 0, |match| or |mismatch| > 63, huge gap costs: classified wide by the engine itself"""
    rng = np.random.default_rng(41)
    jobs = []
    for sc in [(1, -1, 0, -1), (1, -1, -2, 0), (1, -1, 0, 0), (1, -1, 1, -2), (2, -3, -1, 2), (100, -100, -300, -20), (64, -64, -10, -1),
               (1000, -2000, -5000, -1000), (1, -1, -6000, -1), (-1, -1, -2, -1), (3, 5, -4, -2), (16383, -16383, -30000, -2000)]:
        for flag in FLAGS:
            w, q = sv_pair(rng, int(rng.integers(300, 700)), int(rng.integers(20, 90)), ["del", "ins", "tdup", "inv"][flag % 4])
            jobs.append((w, q, flag, *sc))
    check_against_oracle(eng, jobs)


def test_mixed_batch_wide_and_packed(eng):
    """wide and packed jobs in one batch, -both partners across the two kinds of kernels"""
    rng = np.random.default_rng(43)
    jobs = []
    for i in range(24):
        w, q = sv_pair(rng, int(rng.integers(420, 900)), int(rng.integers(30, 120)), ["del", "ins", "tdup", "inv"][i % 4])
        sc = (1, -1, -10, -1) if i % 2 else (90, -70, -200, -15)
        flag = FLAGS[i % len(FLAGS)]
        k = len(jobs)
        jobs.append((w, q, flag, *sc, k + 1))
        jobs.append((w, U.revcom_str(q.decode()).encode(), flag, *sc, k))
    res = eng.align(jobs)
    for i in range(0, len(jobs), 2):
        wa, wb = U.oracle_aligner(*jobs[i][:7]), U.oracle_aligner(*jobs[i + 1][:7])
        a, b = res[i], res[i + 1]
        assert (a["score"], b["score"]) == (wa["score"], wb["score"])
        win, want = (a, wa) if wa["score"] >= wb["score"] else (b, wb)
        lose = b if win is a else a
        assert win["detail"] == 1 and lose["detail"] == 0
        from test_gpu_parity import diff_summary
        assert not diff_summary(win, want), (i, diff_summary(win, want))


def long_pair(rng, len1, len2, mode):
    """a window and a contig of about len2 bases around one event in the middle of the window"""
    w = rand_seq(rng, len1, "ACGT")
    half = len2 // 2
    L = int(rng.integers(150, 400))
    a = len1 // 2 - L // 2
    hl, hr = min(half, a), min(half, len1 - a - L)
    if mode == "del":
        q = w[a - hl:a] + w[a + L:a + L + hr]
    elif mode == "tdup":
        q = w[a + L - min(hl, L):a + L] + w[a:a + hr]
    else:
        q = w[a - hl:a] + U.revcom_str(w[a:a + hr].decode()).encode()
    return w, mutate(rng, q, 0.01, 0.001)


@pytest.mark.parametrize("len1,len2,sc,flag,mode", [
    (12000, 10000, (2, -2, -10, -1), capi.INDEL_FLAG, "del"),        # match 2 with a 10 kb contig: scores to 20 000
    (10000, 8200, (4, -4, -12, -2), capi.INDEL_FLAG, "del"),         # match 4, 8 kb: scores to 32 800 (still exact u16)
    (20000, 20000, (1, -1, -10, -1), capi.INDEL_FLAG, "del"),        # 20 kb x 20 kb at match 1
    (9000, 6000, (2, -3, -8, -2), capi.TDUPLICATION_FLAG, "tdup"),
    (9000, 6000, (3, -3, -9, -1), capi.INVERSION_FLAG, "inv"),
])
def test_long_contigs_beyond_the_16_bit_halves(eng, len1, len2, sc, flag, mode):
    rng = np.random.default_rng(len1 + len2)
    w, q = long_pair(rng, len1, len2, mode)
    check_against_oracle(eng, [(w, q, flag, *sc)])


@pytest.mark.parametrize("flag", [capi.INDEL_FLAG, capi.TDUPLICATION_FLAG, capi.INVL_FLAG])
def test_unsigned_short_wrap(eng, flag):
    """scores beyond 65535: the reference's matrices hold them modulo 65536 and so must we (match 40 on 2.5 kb of near-identity)"""
    rng = np.random.default_rng(7 + flag)
    w = rand_seq(rng, 3000, "ACGT")
    q = mutate(rng, w[200:2700], 0.01, 0.002)
    jobs = [(w, q, flag, 40, -30, -50, -10), (w, q[:1500] + q[1700:], flag, 33, -33, -40, -3)]
    want = [U.oracle_aligner(*j) for j in jobs]
    assert max(x["score"] for x in want) <= 65535          # what comes out is a u16 ...
    check_against_oracle(eng, jobs)                          # ... and identical, wrap and all


def test_status_codes(eng):
    A = b"ACGTACGTAC"
    res = eng.align([(A, A, capi.INDEL_FLAG, 100, -1, -10, -1),                 # wide now, not a range error
                     (A, A, capi.INDEL_FLAG, 20000, -1, -10, -1),               # 2*match does not fit the 16-bit score tables
                     (A, A, capi.INDEL_FLAG | capi.TDUPLICATION_FLAG, 1, -1, -10, -1),
                     (A, A, capi.INVERSION_FLAG | capi.INDEL_FLAG, 1, -1, -10, -1),
                     (A, A, 0x40, 1, -1, -10, -1)])
    assert [r["status"] for r in res] == [capi.AGE_OK, capi.AGE_ERR_RANGE, capi.AGE_ERR_ARG, capi.AGE_OK, capi.AGE_NOT_ALIGNED]
    want = U.oracle_aligner(A, A, capi.INVERSION_FLAG, 1, -1, -10, -1)
    assert res[3]["score"] == want["score"] and res[3]["bp"] == want["bp"]
