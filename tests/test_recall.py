"""Variant re-calling from the excised region: age_recall_variant / age_format_vcf_age against

  * the product's own AlignResult path (age_describe_alignment, pinned to 370 dumps of the reference aligner by
    test_align_result.py): identities, strand, fragment coordinates and confidence intervals must agree, and
  * the variants the unmodified AsmvarDetect sources re-call (tests/golden/recall_cases.json, 85 cases over
    DEL / INS / SNP / MNP / REPLACEMENT): type, coordinates, cipos / ciend, the good-alignment rule, the AGE= field.

CPU part: alignments from the oracle, age_counts by their definitions in Python (ageutil.counts_by_definition).  The GPU
part (test_gpu_recall.py) checks that the device computes exactly these counts."""
import json

import numpy as np
import pytest

import ageutil as U
from asmvar_b200 import capi

CASES = U.load_cases()
GOLD = json.loads((U.GOLDEN / "align_result_cases.json").read_text())["results"]
BY_ID = {c["id"]: c for c in CASES}


def recall_from_summary(s1, s2, summary):
    r, keep = U.summary_to_result(summary)
    U.set_counts(r, U.counts_by_definition(s1["seq"], s2["seq"], summary))
    v1 = U.make_view(s1["seq"].encode(), s1["name"], s1["start"], s1["reverse"])
    v2 = U.make_view(s2["seq"].encode(), s2["name"], s2["start"], s2["reverse"])
    return capi.recall_variant(v1, v2, r), capi.describe_alignment(v1, v2, r)


@pytest.mark.parametrize("g", GOLD, ids=[g["id"] for g in GOLD])
def test_recall_agrees_with_align_result(g):
    case = BY_ID[g["case"]]
    s1 = case["s1"]
    s2 = U.revcom_view(case["s2"]) if g["rc"] else case["s2"]
    summary = U.oracle_aligner(s1["seq"].encode(), s2["seq"].encode(), case["flag"], case["match"], case["mismatch"], case["go"], case["ge"])
    if summary["status"] != 0:
        return
    v, d = recall_from_summary(s1, s2, summary)
    assert v["identity"] == d["identity"] and v["strand"] == d["strand"]
    assert v["frag"] == [[m["start1"], m["end1"], m["start2"], m["end2"]] for m in d["map"]]
    for k in ("ci_start1", "ci_end1", "ci_start2", "ci_end2"):
        assert v[k] == d[k], (g["id"], k)
    assert v["homo_run"] == d["homo_run"][0]
    assert v["has_variant"] == (len(d["map"]) == 2)
    # the counts by definition reproduce what the string-based path counts
    c = U.counts_by_definition(s1["seq"], s2["seq"], summary)
    sides = [k for k, name in enumerate(("left", "right")) if summary[name]]
    assert [c["n_aligned"][k] for k in sides] == [m["n_aligned"] for m in d["map"]]
    assert [c["n_identic"][k] for k in sides] == [m["n_identic"] for m in d["map"]]
    assert [c["n_gap"][k] for k in sides] == [m["n_gap"] for m in d["map"]]
    if len(d["map"]) == 2:
        assert [c["homo_at1"][0], c["homo_at2"][0], c["homo_in"][0], c["homo_in"][1], c["homo_out"][0], c["homo_out"][1]] == d["homo_run"]


def recall_cases():
    d = json.loads((U.GOLDEN / "recall_cases.json").read_text())
    return [c for c in d["cases"] if "regen" not in c]


@pytest.mark.parametrize("c", recall_cases(), ids=[c["id"] for c in recall_cases()])
def test_recalled_variant_matches_asmvardetect(c):
    """the flow of VarUnit::ReAlignAndReCallVar (VarUnit.cpp:215-336) around one candidate, then VarReCall"""
    exp = c["expected"]
    if exp is None or not exp["success"]:
        return
    ts, te = max(1, c["tstart"] - c["extend"]), min(len(c["tseq"]), c["tend"] + c["extend"])     # ExtendVU
    qs, qe = max(1, c["qstart"] - c["extend"]), min(len(c["qseq"]), c["qend"] + c["extend"])
    s1 = {"name": c["tid"], "start": ts, "reverse": False, "seq": c["tseq"][ts - 1:te]}
    s2 = {"name": c["qid"], "start": qs, "reverse": False, "seq": c["qseq"][qs - 1:qe]}
    sc = (c["match"], c["mismatch"], c["go"], c["ge"])
    variants = [s2] + ([U.revcom_view(s2)] if c["both"] else [])
    res = [U.oracle_aligner(s1["seq"].encode(), v["seq"].encode(), c["flag"], *sc) for v in variants]
    pick = 0 if len(res) == 1 or res[0]["score"] >= res[1]["score"] else 1
    if res[pick]["used_aux"]:
        assert exp["vus"] == []               # SetAlignResult printed the auxiliary record and left AlignResult empty
        return
    v, _ = recall_from_summary(s1, variants[pick], res[pick])
    if not exp["vus"]:
        assert not v["has_variant"]
        return
    w = exp["vus"][0]
    assert v["has_variant"] and v["type_name"] == w["type"] and v["strand"] == w["strand"]
    assert [v["target"][0], v["target"][1], v["query"][0], v["query"][1]] == [w["tstart"], w["tend"], w["qstart"], w["qend"]]
    assert v["cipos"] == w["cipos"] and v["ciend"] == w["ciend"] and v["homo_run"] == w["homoRun"]
    assert v["identity"] == w["identity"]
    good = v["good_align"] and c["mismap"] < 0.01
    assert int(good) == w["isGoodReAlign"]
    i = w["identity"]
    assert v["AGE=T" if good else "AGE=F"] == f"{'T' if good else 'F'},{w['strand']},{i[0][0]},{i[0][1]},{i[1][0]},{i[1][1]},{i[2][0]},{i[2][1]}"


def test_age_field_without_excised_region():
    s = "ACGTTGCAAGGCTTAACCGGTTAA"
    summary = U.oracle_aligner(s.encode(), s.encode(), capi.INDEL_FLAG, 1, -1, -10, -1)
    v, _ = recall_from_summary({"name": "t", "start": 1, "reverse": False, "seq": s}, {"name": "q", "start": 1, "reverse": False, "seq": s}, summary)
    assert not v["has_variant"] and v["type_name"] == "" and v["AGE=T"] == "." and v["identity"] == [[len(s), 100]]
