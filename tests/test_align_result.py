"""AsmvarDetect's structured result (SetAlignResult / AlignResult, src/AsmvarDetect/AGEaligner.cpp:179-332) against
the dumps of the unmodified reference aligner (tests/golden/align_result_cases.json, made by oracle/make_golden_detect.py).

CPU part: oracle alignment -> the PRODUCT's age_describe_alignment.  GPU part (test_shim.py) runs the same cases through
include/age_shim.hh."""
import json

import pytest

import ageutil as U
from asmvar_b200 import capi

GOLD = json.loads((U.GOLDEN / "align_result_cases.json").read_text())["results"]
CASES = {c["id"]: c for c in U.load_cases()}
FIELDS = ["score", "strand", "is_alternative_align", "identity", "ci_start1", "ci_end1", "ci_start2", "ci_end2", "homo_run"]


def inputs(g):
    case = CASES[g["case"]]
    s2 = U.revcom_view(case["s2"]) if g["rc"] else case["s2"]
    return case, case["s1"], s2


def compare(got: dict, g: dict):
    for k in FIELDS:
        assert got[k] == g[k], (g["id"], k, got[k], g[k])
    assert len(got["map"]) == len(g["map"])
    for a, b in zip(got["map"], g["map"]):
        for k in ("start1", "end1", "start2", "end2", "seq1", "seq2", "info"):
            assert a[k] == b[k], (g["id"], k)


@pytest.mark.parametrize("g", GOLD, ids=[g["id"] for g in GOLD])
def test_describe_alignment_matches_reference_align_result(g):
    case, s1, s2 = inputs(g)
    summary = U.oracle_aligner(s1["seq"].encode(), s2["seq"].encode(), case["flag"], case["match"], case["mismatch"], case["go"], case["ge"])
    assert (summary["status"] == 0) == bool(g["aligned"])
    if not g["aligned"]:
        return
    if g["printed_bytes"]:
        # the auxiliary INVR aligner won: the reference prints its record and leaves AlignResult at its defaults
        assert summary["used_aux"] == 1 and g["score"] == -1 and g["map"] == []
        return
    assert summary["used_aux"] == 0
    r, keep = U.summary_to_result(summary)
    b1, b2 = s1["seq"].encode(), s2["seq"].encode()
    got = capi.describe_alignment(U.make_view(b1, s1["name"], s1["start"], s1["reverse"]), U.make_view(b2, s2["name"], s2["start"], s2["reverse"]), r)
    compare(got, g)
    assert g["id1"] == s1["name"] and g["id2"] == s2["name"]


def test_golden_inventory():
    assert len(GOLD) >= 300
    assert any(g["printed_bytes"] for g in GOLD) and any(len(g.get("map", [])) == 2 for g in GOLD) and any(len(g.get("map", [])) == 1 for g in GOLD)
    assert any(g.get("strand") == "-" for g in GOLD) and any(g.get("is_alternative_align") for g in GOLD)
    assert any(g.get("ci_start2", [0, 0]) != [0, 0] for g in GOLD)
