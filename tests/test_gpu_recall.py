"""The numbers the device leaves in age_result::counts (column counts of the two fragments, the runs around the excised
region) against their definitions in plain Python, and the variants re-called from them against the unmodified
AsmvarDetect sources (tests/golden/recall_cases.json) -- SURVEY 8f-3 on the GPU path."""
import json

import numpy as np
import pytest

import ageutil as U
from asmvar_b200 import capi
from test_gpu_parity import CASES, FLAGS, case_jobs, random_jobs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from asmvar_b200 import engine
    with engine.Engine() as e:
        yield e


def check_counts(eng, jobs):
    res = eng.align(jobs)
    bad = []
    for i, (j, r) in enumerate(zip(jobs, res)):
        if r["status"] != 0 or not r["detail"]:
            continue
        want = U.counts_by_definition(j[0].decode("latin-1"), j[1].decode("latin-1"), r)
        if r["counts"] != want:
            bad.append((i, r["flag"], {k: (r["counts"][k], want[k]) for k in want if r["counts"][k] != want[k]}))
    assert not bad, f"{len(bad)} jobs with other counts than the definition: {bad[:5]}"


def test_counts_on_the_golden_batch(eng):
    jobs = []
    for c in CASES:
        jobs += case_jobs(c, len(jobs), True)
    check_counts(eng, jobs)


@pytest.mark.parametrize("knob", [{}, {"AGE_FORCE_WIDE": "1"}])
def test_counts_on_random_jobs(eng, knob, monkeypatch):
    for k, v in knob.items():
        monkeypatch.setenv(k, v)
    check_counts(eng, random_jobs(51, 160, (60, 500), (15, 90), ["ACGT", "ACGT", "ACGTN", "ACGTacgtNnU"]))
    check_counts(eng, random_jobs(52, 24, (900, 2200), (200, 420), ["ACGT", "ACGTN"]))


def test_counts_on_repeats(eng):
    """tandem repeats and homopolymers around the event: long runs at, outside and inside the breakpoints"""
    rng = np.random.default_rng(8)
    jobs = []
    for i in range(60):
        unit = "".join(rng.choice(list("ACGT"), size=int(rng.integers(1, 9))))
        rep = unit * int(rng.integers(3, 40))
        left = "".join(rng.choice(list("ACGT"), size=int(rng.integers(40, 120))))
        right = "".join(rng.choice(list("ACGT"), size=int(rng.integers(40, 120))))
        w = left + rep + right
        cut = len(unit) * int(rng.integers(1, max(2, len(rep) // len(unit))))
        q = left + rep[cut:] + right                       # a deletion of whole repeat units: the breakpoint can slide
        flag = FLAGS[i % len(FLAGS)]
        jobs.append((w.encode(), q.encode(), flag, 1, -1, -10, -1))
        jobs.append((q.encode(), w.encode(), capi.INDEL_FLAG, 1, -1, -10, -1))       # the same as an insertion
    check_counts(eng, jobs)


def test_recalled_variants_match_asmvardetect(eng):
    d = json.loads((U.GOLDEN / "recall_cases.json").read_text())
    cases = [c for c in d["cases"] if "regen" not in c and c["expected"] and c["expected"]["success"]]
    jobs, meta = [], []
    for c in cases:
        ts, te = max(1, c["tstart"] - c["extend"]), min(len(c["tseq"]), c["tend"] + c["extend"])
        qs, qe = max(1, c["qstart"] - c["extend"]), min(len(c["qseq"]), c["qend"] + c["extend"])
        s1 = {"name": c["tid"], "start": ts, "reverse": False, "seq": c["tseq"][ts - 1:te]}
        s2 = {"name": c["qid"], "start": qs, "reverse": False, "seq": c["qseq"][qs - 1:qe]}
        variants = [s2] + ([U.revcom_view(s2)] if c["both"] else [])
        k = len(jobs)
        for n, v in enumerate(variants):
            partner = -1 if len(variants) == 1 else (k + 1 if n == 0 else k)
            jobs.append((s1["seq"].encode(), v["seq"].encode(), c["flag"], c["match"], c["mismatch"], c["go"], c["ge"], partner))
        meta.append((c, s1, variants, k))
    arr, keep, n = eng.make_jobs(jobs)
    res = eng.align_raw(arr, n)
    n_var = 0
    for c, s1, variants, k in meta:
        pick = 0 if len(variants) == 1 or res[k].score >= res[k + 1].score else 1
        r, s2 = res[k + pick], variants[pick]
        exp = c["expected"]["vus"]
        if r.used_aux:
            assert exp == []
            continue
        v = capi.recall_variant(U.make_view(s1["seq"].encode(), s1["name"], s1["start"], False), U.make_view(s2["seq"].encode(), s2["name"], s2["start"], s2["reverse"]), r)
        if not exp:
            assert not v["has_variant"], c["id"]
            continue
        w = exp[0]
        n_var += 1
        assert v["type_name"] == w["type"] and v["strand"] == w["strand"], c["id"]
        assert [v["target"][0], v["target"][1], v["query"][0], v["query"][1]] == [w["tstart"], w["tend"], w["qstart"], w["qend"]], c["id"]
        assert v["cipos"] == w["cipos"] and v["ciend"] == w["ciend"] and v["homo_run"] == w["homoRun"] and v["identity"] == w["identity"], c["id"]
        assert int(v["good_align"] and c["mismap"] < 0.01) == w["isGoodReAlign"], c["id"]
    assert n_var >= 60
