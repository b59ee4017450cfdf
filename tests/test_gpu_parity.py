"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the reference's golden records.
Integer work, so everything is compared bit-exactly: score, breakpoint list, diagonal blocks and the `.age` text."""
import numpy as np
import pytest

import ageutil as U
from asmvar_b200 import capi

pytestmark = pytest.mark.gpu

CASES = U.load_cases()
KEYS = ("status", "score", "flag", "used_aux", "main_score", "aux_score", "bp", "alt_index", "left", "right", "alt_left", "alt_right")


@pytest.fixture(scope="module")
def eng():
    from asmvar_b200 import engine
    with engine.Engine() as e:
        yield e


def case_jobs(case, base=0, partners=True):
    """jobs of one candidate; with `partners` the two -both jobs name each other (age_job.partner)"""
    s1, s2 = case["s1"]["seq"].encode(), case["s2"]["seq"].encode()
    sc = (case["match"], case["mismatch"], case["go"], case["ge"])
    if not case["both"]:
        return [(s1, s2, case["flag"], *sc, -1)]
    p = (base + 1, base) if partners else (-1, -1)
    return [(s1, s2, case["flag"], *sc, p[0]), (s1, U.revcom_str(case["s2"]["seq"]).encode(), case["flag"], *sc, p[1])]


def diff_summary(got, want):
    return [k for k in KEYS if k in want and got.get(k) != want.get(k)]


@pytest.mark.parametrize("partners", [True, False])
def test_golden_records_one_batch(eng, partners):
    """All 247 golden records (the reference's 15 known-answer cases, Test1, Test2, 220 reference-generated
    random cases) in ONE batch: exercises pairing, mixed shapes and modes, with and without the -both
    partner selection done inside the library."""
    jobs, owner = [], []
    for ci, c in enumerate(CASES):
        for j in case_jobs(c, len(jobs), partners):
            jobs.append(j); owner.append(ci)
    res = eng.align(jobs)
    if partners:
        for k, j in enumerate(jobs):
            if j[7] > k and j[2] != capi.INVERSION_FLAG:      # exactly one job of a -both couple carries breakpoints, the one age_align prints
                a, b = res[k], res[j[7]]
                assert a["detail"] + b["detail"] == 1
                assert a["detail"] == (1 if a["score"] >= b["score"] else 0)
    by_case = {}
    for r, ci in zip(res, owner):
        by_case.setdefault(ci, []).append(r)
    bad = []
    for ci, c in enumerate(CASES):
        it = iter(by_case[ci])
        got = U.candidate_record(c, lambda *a: next(it))
        if got.rstrip("\n") != c["expected"].rstrip("\n"):
            bad.append(c["id"])
    assert not bad, f"{len(bad)} records differ from the reference: {bad[:20]}"


def rand_seq(rng, n, alphabet):
    return "".join(rng.choice(list(alphabet), size=n)).encode()


def mutate(rng, s, sub=0.03, indel=0.01):
    out = []
    for ch in s.decode():
        x = rng.random()
        if x < sub:
            out.append(rng.choice(list("ACGT")))
        elif x < sub + indel:
            continue
        elif x < sub + 2 * indel:
            out.append(ch); out.append(rng.choice(list("ACGT")))
        else:
            out.append(ch)
    return "".join(out).encode()


def sv_pair(rng, len1, flank, mode, alphabet="ACGT"):
    """A reference window and a contig carrying one structural event, so the split alignment is non-trivial."""
    w = rand_seq(rng, len1, alphabet)
    a = int(rng.integers(flank, len1 - flank - 10))
    L = int(rng.integers(5, max(6, len1 - a - flank)))
    left, right = w[a - flank:a], w[a + L:a + L + flank]
    if mode == "del":
        q = left + right
    elif mode == "ins":
        q = w[a - flank:a] + rand_seq(rng, int(rng.integers(3, 60)), "ACGT") + w[a:a + flank]
    elif mode == "tdup":
        q = w[a - flank:a + L][-flank:] + w[a:a + flank]
    else:  # inversion
        q = left + U.revcom_str(w[a:a + flank].decode()).encode()
    q = mutate(rng, q)
    if rng.random() < 0.3:
        q = U.revcom_str(q.decode()).encode()
    return w, q


SCORINGS = [(1, -1, -10, -1), (1, -2, -2, -1), (2, -3, -5, -2), (3, -1, -4, -4), (1, -1, -1, -3)]
FLAGS = [capi.INDEL_FLAG, capi.TDUPLICATION_FLAG, capi.INVERSION_FLAG, capi.INVL_FLAG, capi.INVR_FLAG]


def random_jobs(seed, n, len1_range, flank_range, alphabets):
    rng = np.random.default_rng(seed)
    jobs = []
    for i in range(n):
        len1 = int(rng.integers(*len1_range))
        flank = int(rng.integers(*flank_range))
        flank = min(flank, (len1 - 20) // 3)
        mode = ["del", "ins", "tdup", "inv"][i % 4]
        w, q = sv_pair(rng, len1, flank, mode, alphabets[i % len(alphabets)])
        flag = FLAGS[int(rng.integers(0, len(FLAGS)))] if i % 3 else capi.INDEL_FLAG
        jobs.append((w, q, flag, *SCORINGS[int(rng.integers(0, len(SCORINGS)))]))
    return jobs


def check_against_oracle(eng, jobs):
    res = eng.align(jobs)
    bad = []
    for i, (j, r) in enumerate(zip(jobs, res)):
        want = U.oracle_aligner(*j)
        d = diff_summary(r, want)
        if d:
            bad.append((i, j[2], len(j[0]), len(j[1]), d, r.get("score"), want.get("score")))
    assert not bad, f"{len(bad)}/{len(jobs)} jobs differ from the oracle; first: {bad[:8]}"


def test_random_small_all_modes(eng):
    check_against_oracle(eng, random_jobs(11, 160, (60, 400), (15, 80), ["ACGT", "ACGT", "ACGTN", "ACGTacgtNnU"]))


def test_random_multi_strip(eng):
    """contigs longer than one 256-row strip: the strip boundary rows and the masked last strip are exercised"""
    check_against_oracle(eng, random_jobs(12, 40, (900, 2200), (200, 420), ["ACGT", "ACGTN"]))


def test_homopolymer_and_degenerate(eng):
    jobs = []
    for flag in FLAGS:
        jobs.append((b"A" * 90, b"A" * 40, flag, 1, -1, -10, -1))           # every cell ties
        jobs.append((b"N" * 50, b"N" * 70, flag, 1, -1, -10, -1))           # all scores zero
        jobs.append((b"ACGT" * 30, b"TTTT" + b"ACGT" * 10, flag, 1, -2, -2, -1))
        jobs.append((b"A", b"A", flag, 1, -1, -10, -1))
        jobs.append((b"A", b"C", flag, 1, -1, -10, -1))
    check_against_oracle(eng, jobs)


def gapped_copy(rng, s, every):
    """`s` with one base dropped about every `every` bases: an alignment of many short diagonal blocks"""
    out, i = bytearray(), 0
    while i < len(s):
        n = int(rng.integers(max(2, every - 3), every + 4))
        out += s[i:i + n]
        i += n + 1
    return bytes(out)


def test_many_blocks_per_fragment(eng):
    """alignments of 100+ diagonal blocks per fragment (cheap gaps, a base missing every ~9 bases of the flanks): more than the
    blocks kernel stages in shared memory, so the second traversal that writes straight into the pool runs as well"""
    rng = np.random.default_rng(21)
    jobs = []
    for flag, L in ((capi.INDEL_FLAG, 700), (capi.INDEL_FLAG, 40), (capi.TDUPLICATION_FLAG, 500)):
        w = rand_seq(rng, 3600, "ACGT")
        a = 1300
        if flag == capi.INDEL_FLAG:
            q = gapped_copy(rng, w[a - 1000:a], 9) + gapped_copy(rng, w[a + L:a + L + 1000], 9)
        else:
            q = gapped_copy(rng, w[a + L - 1000:a + L], 9) + gapped_copy(rng, w[a:a + 1000], 9)
        jobs.append((w, q, flag, 5, -4, -2, -1))
    res = eng.align(jobs)
    assert max(len(r["left"]) for r in res) > 64 and max(len(r["right"]) for r in res) > 64, [(len(r["left"]), len(r["right"])) for r in res]
    check_against_oracle(eng, jobs)


def test_long_runs_and_wide_ranges(eng):
    """walks along the whole excised stretch (3 kb deletion, 1.2 kb insertion: straight runs of the maxima planes read 256
    cells per pass, look-ahead stops at windows that were never swept) and split columns several cells wide (the last bases
    of the left flank repeat behind the insertion, so that many rows tie in several columns)"""
    rng = np.random.default_rng(22)
    jobs = []
    w = rand_seq(rng, 6000, "ACGT")
    jobs.append((w, w[800:1400] + w[4400:5000], capi.INDEL_FLAG, 1, -1, -10, -1))                                # deletion of 3 kb
    jobs.append((w, w[2000:2600] + rand_seq(rng, 1200, "ACGT") + w[2600:3200], capi.INDEL_FLAG, 1, -1, -10, -1))  # insertion of 1.2 kb
    rep = w[2588:2600]                                                                                          # micro-homology at the junction
    jobs.append((w, w[2000:2600] + rep * 3 + rand_seq(rng, 900, "ACGT") + rep + w[2600:3200], capi.INDEL_FLAG, 1, -1, -10, -1))
    jobs.append((w, mutate(rng, w[1000:1700] + rand_seq(rng, 700, "ACGT") + w[1700:2400], 0.02, 0.004), capi.INDEL_FLAG, 1, -1, -10, -1))
    jobs.append((w, w[3000:3700] + w[1500:2200], capi.TDUPLICATION_FLAG, 1, -1, -10, -1))                          # tandem duplication junction
    jobs.append((w, w[1000:1800] + U.revcom_str(w[1800:2600].decode()).encode(), capi.INVERSION_FLAG, 1, -1, -10, -1))
    check_against_oracle(eng, jobs)


def test_not_aligned_and_range(eng):
    res = eng.align([(b"", b"ACGT", capi.INDEL_FLAG, 1, -1, -10, -1), (b"ACGT", b"ACGT", 0, 1, -1, -10, -1),
                     (b"ACGT", b"ACGT", capi.INDEL_FLAG, 20000, -1, -10, -1), (b"ACGT", b"ACGT", capi.INDEL_FLAG, 1, -1, -10, -1)])
    assert [r["status"] for r in res] == [capi.AGE_NOT_ALIGNED, capi.AGE_NOT_ALIGNED, capi.AGE_ERR_RANGE, capi.AGE_OK]


def test_batch_order_independent(eng):
    """the same jobs in another order (hence other pairings) give the same per-job results"""
    jobs = random_jobs(13, 48, (80, 300), (20, 60), ["ACGT", "ACGTN"])
    a = eng.align(jobs)
    perm = list(np.random.default_rng(5).permutation(len(jobs)))
    b = eng.align([jobs[i] for i in perm])
    for k, i in enumerate(perm):
        assert not diff_summary(b[k], a[i])


def test_two_devices_same_results(eng):
    """age_create over two GPUs shards the pairs (no collective, host gather): per-job results are identical"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from asmvar_b200 import engine
    jobs = random_jobs(21, 96, (200, 1200), (40, 300), ["ACGT", "ACGTN"])
    a = eng.align(jobs)
    with engine.Engine([0, 1]) as e2:
        b = e2.align(jobs)
    for x, y in zip(a, b):
        assert not diff_summary(y, x)


@pytest.mark.parametrize("knob", [{"AGE_LONG_NW": "1"}, {"AGE_LONG_NW": "16"}, {"AGE_DFMT": "raw"}, {"AGE_DFMT": "raw", "AGE_LONG_NW": "1"},
                                  {"AGE_QUEUE_CAP": "4"}, {"AGE_QUEUE_CAP": "4", "AGE_LONG_NW": "1"}, {"AGE_TRACE_POOL_GRANULES": "400"},
                                  {"AGE_TRACE_POOL_GRANULES": "400", "AGE_QUEUE_CAP": "4"}, {"AGE_LONG_NW": "8", "AGE_LONG_CS": "2"},
                                  {"AGE_LONG_NW": "16", "AGE_LONG_CS": "4"}])
def test_golden_records_other_kernel_paths(eng, knob, monkeypatch):
    """The same golden batch through the code paths a small batch does not take by itself: one warp per pair
    (the kernel large batches use), 16 warps per pair (long alignments), a work list too short for the
    pre-sweeps, so that pass 2 re-sweeps nearly every window on demand, the 16-bit storage format of the
    forward-maxima plane (by default only scorings with match or mismatch above 15 use it), and a trace pool far too
    small for the batch: the engine notices the overflow and runs the batch again with a larger pool; thread-block
    clusters of 2 / 4 CTAs sharing one pair (what a few dozen very long pairs get)."""
    for k, v in knob.items():
        monkeypatch.setenv(k, v)
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
    assert not bad, f"{knob}: {len(bad)} records differ from the reference: {bad[:20]}"


def test_geometry_boundaries(eng):
    """Shapes that sit on the seams of the kernel geometry: strips of 256 rows, blocks of 4 columns, windows of
    32 steps (128 columns), fewer blocks than lanes, single rows / columns."""
    rng = np.random.default_rng(99)
    len2s = [1, 2, 7, 8, 9, 255, 256, 257, 511, 512, 513, 768]
    len1s = [1, 2, 3, 4, 5, 31, 32, 33, 124, 127, 128, 129, 132, 257, 1023, 1025]
    jobs = []
    for i, l2 in enumerate(len2s):
        for k, l1 in enumerate(len1s):
            if (i + k) % 3:               # a third of the grid is enough for one batch
                continue
            w = rand_seq(rng, l1, "ACGT")
            # a contig that shares material with the window (so that scores are not trivial) cut or padded to l2
            base = (w * (l2 // max(l1, 1) + 2))[: l2 + 7]
            q = mutate(rng, base, 0.05, 0.02)[:l2]
            if len(q) < l2:
                q = q + rand_seq(rng, l2 - len(q), "ACGT")
            flag = FLAGS[(i + 2 * k) % len(FLAGS)]
            jobs.append((w, q, flag, *SCORINGS[(i + k) % len(SCORINGS)]))
    check_against_oracle(eng, jobs)


def test_scoring_corners(eng):
    """gap extension dearer than gap opening (the flipped tag branch), large match scores, equal open / extend"""
    rng = np.random.default_rng(5)
    jobs = []
    # (15, ..) is the largest step the 8-bit row offsets of the forward-maxima plane hold, (16, ..) the first scoring on the
    # 16-bit planes; (2, 9, ..) rewards mismatches more than matches (the offset bound is max(match, mismatch))
    for sc in [(1, -1, -1, -3), (5, -4, -3, -3), (63, -63, -40, -2), (2, -1, -2, -7), (1, -3, -100, -1), (7, -2, -1, -1),
               (15, -4, -3, -3), (16, -4, -3, -3), (2, 9, -3, -2), (15, -15, -1, -1)]:
        for flag in FLAGS:
            w, q = sv_pair(rng, int(rng.integers(120, 420)), int(rng.integers(20, 60)), ["del", "ins", "tdup", "inv"][flag % 4])
            jobs.append((w, q, flag, *sc))
    check_against_oracle(eng, jobs)
