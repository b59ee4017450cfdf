"""Test helpers: ctypes binding of the CPU oracle (oracle/libage_oracle.so) and glue that feeds
oracle results through the PRODUCT formatter (age_format_alignment), so that on a machine without a
GPU the `.age` text path is still checked against the reference's golden records.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may touch oracle/.
"""
from __future__ import annotations

import ctypes as C
import json
from pathlib import Path

from asmvar_b200 import capi

ROOT = Path(__file__).resolve().parent.parent
ORACLE_LIB = ROOT / "oracle" / "libage_oracle.so"
GOLDEN = ROOT / "tests" / "golden"

COMP = {"a": "t", "c": "g", "t": "a", "g": "c", "A": "T", "C": "G", "T": "A", "G": "C", "u": "a", "U": "A"}


def revcom_str(s: str) -> str:
    return "".join(COMP.get(c, c) for c in reversed(s))


class OrcBlock(C.Structure):
    _fields_ = [("start1", C.c_int32), ("start2", C.c_int32), ("length", C.c_int32)]


class OrcResult(C.Structure):
    _fields_ = [
        ("score", C.c_int32), ("n_bp", C.c_int32), ("bp", (C.c_int32 * 4) * 100), ("alt_index", C.c_int32),
        ("n_left", C.c_int32), ("n_right", C.c_int32), ("n_alt_left", C.c_int32), ("n_alt_right", C.c_int32),
        ("left", C.POINTER(OrcBlock)), ("right", C.POINTER(OrcBlock)),
        ("alt_left", C.POINTER(OrcBlock)), ("alt_right", C.POINTER(OrcBlock)),
    ]


_orc = None


def oracle():
    global _orc
    if _orc is None:
        lib = C.CDLL(str(ORACLE_LIB))
        lib.orc_align.restype = C.c_int
        lib.orc_align.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(OrcResult)]
        lib.orc_free_result.argtypes = [C.POINTER(OrcResult)]
        lib.orc_score.restype = C.c_int
        lib.orc_score.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        _orc = lib
    return _orc


def _blocks(p, n):
    return [(p[i].start1, p[i].start2, p[i].length) for i in range(n)]


def oracle_single(seq1: bytes, seq2: bytes, flag: int, match, mismatch, go, ge):
    """One single-mode aligner -> plain dict (None if align() would return false)."""
    r = OrcResult()
    rc = oracle().orc_align(seq1, len(seq1), seq2, len(seq2), flag, match, mismatch, go, ge, C.byref(r))
    if rc != 0:
        return None
    d = {"score": r.score, "bp": [tuple(r.bp[i]) for i in range(r.n_bp)], "alt_index": r.alt_index,
         "left": _blocks(r.left, r.n_left), "right": _blocks(r.right, r.n_right),
         "alt_left": _blocks(r.alt_left, r.n_alt_left), "alt_right": _blocks(r.alt_right, r.n_alt_right), "flag": flag}
    oracle().orc_free_result(C.byref(r))
    return d


def oracle_aligner(seq1: bytes, seq2: bytes, flag: int, match, mismatch, go, ge):
    """AGEaligner::align + score() + aux selection (AGEaligner.cpp:72-77,136-154,181-184) on the oracle.
    Returns a dict shaped like capi.AgeResult.summary()."""
    if flag & capi.INVERSION_FLAG:
        main = oracle_single(seq1, seq2, capi.INVL_FLAG, match, mismatch, go, ge)
        aux = oracle_single(seq1, seq2, capi.INVR_FLAG, match, mismatch, go, ge) if main else None
    else:
        main = oracle_single(seq1, seq2, flag, match, mismatch, go, ge)
        aux = None
    if main is None:
        return {"status": capi.AGE_NOT_ALIGNED}
    used_aux = bool(aux and aux["score"] > main["score"])
    p = aux if used_aux else main
    return {"status": 0, "score": p["score"], "flag": p["flag"], "used_aux": int(used_aux),
            "main_score": main["score"], "aux_score": aux["score"] if aux else -1,
            "bp": p["bp"], "alt_index": p["alt_index"], "left": p["left"], "right": p["right"],
            "alt_left": p["alt_left"] if p["alt_index"] >= 1 else [], "alt_right": p["alt_right"] if p["alt_index"] >= 1 else []}


def summary_to_result(d: dict):
    """dict -> (capi.AgeResult, keepalive) so the product formatter can print an oracle result."""
    r = capi.AgeResult()
    keep = []
    r.status, r.score, r.flag, r.used_aux = d["status"], d["score"], d["flag"], d["used_aux"]
    r.detail = 1
    r.main_score, r.aux_score = d["main_score"], d["aux_score"]
    r.n_bp = len(d["bp"])
    for i, b in enumerate(d["bp"]):
        for k in range(4):
            r.bp[i][k] = b[k]
    r.alt_index = d["alt_index"]
    for name in ("left", "right", "alt_left", "alt_right"):
        arr = (capi.AgeBlock * max(1, len(d[name])))()
        for i, b in enumerate(d[name]):
            arr[i].start1, arr[i].start2, arr[i].length = b
        keep.append(arr)
        setattr(r, "n_" + name, len(d[name]))
        setattr(r, name, C.cast(arr, C.POINTER(capi.AgeBlock)))
    return r, keep


# ---- age_counts by the definitions (what the device computes from the block lists; tests/test_recall.py) ---------------
def _same(a, b):
    a, b = a.upper(), b.upper()
    return a == b and a in "ACGTU"


def _ch(s, i):
    return s[i] if 0 <= i < len(s) else "\0"


def run_at(s, bp1, bp2):
    """calcIdentity (AGEaligner.cpp:429-447): (run, left extent <= 0, right extent >= 0)"""
    if bp1 >= bp2:
        return [0, 0, 0]
    d, down, up = bp2 - bp1, 0, 0
    i = bp1 - 1
    while i >= 0 and _same(_ch(s, i), _ch(s, i + d)):
        down += 1; i -= 1
    i = bp1
    while i < len(s) and _same(_ch(s, i), _ch(s, i + d)):
        up += 1; i += 1
    return [up + down, -down, up]


def run_outside(s, bp1, bp2):
    if bp1 >= bp2:
        return 0
    for m in range(min(bp1 + 1, len(s) - bp2), 0, -1):
        if all(_same(_ch(s, bp1 - m + 1 + k), _ch(s, bp2 + k)) for k in range(m)):
            return m
    return 0


def run_inside(s, bp1, bp2):
    if bp1 >= bp2:
        return 0
    for m in range((bp2 - bp1 + 1) // 2, 0, -1):
        if all(_same(_ch(s, bp1 + k), _ch(s, bp2 - m + 1 + k)) for k in range(m)):
            return m
    return 0


def counts_by_definition(seq1: str, seq2: str, summary: dict) -> dict:
    """age_counts of a result summary, in plain Python: columns of the two fragments from their diagonal blocks, and the
    three runs around the excised region in positions of the sequences as passed."""
    flag, len1 = summary["flag"], len(seq1)
    rc = revcom_str(seq1)
    out = {"n_aligned": [0, 0], "n_identic": [0, 0], "n_gap": [0, 0], "homo_at1": [0, 0, 0], "homo_at2": [0, 0, 0], "homo_out": [0, 0], "homo_in": [0, 0]}
    for k, name in enumerate(("left", "right")):
        bl = summary[name]
        on_rc = (flag & capi.INVL_FLAG and k == 1) or (flag & capi.INVR_FLAG and k == 0)
        q1 = rc if on_rc else seq1
        for i, (s1, s2, n) in enumerate(bl):
            if i + 1 < len(bl):
                g = (bl[i + 1][0] - (s1 + n)) + (bl[i + 1][1] - (s2 + n))
                out["n_aligned"][k] += g; out["n_gap"][k] += g
            for j in range(n):
                a, b = q1[s1 - 1 + j], seq2[s2 - 1 + j]
                out["n_aligned"][k] += 1
                out["n_identic"][k] += int(_same(a, b))
                out["n_gap"][k] += int(a in " -" or b in " -")
    if summary["left"] and summary["right"]:
        le, rs = summary["left"][-1], summary["right"][0]
        e1, e2 = le[0] + le[2] - 1, le[1] + le[2] - 1
        A = len1 - e1 if flag & capi.INVR_FLAG else e1 - 1
        B = len1 - rs[0] if flag & capi.INVL_FLAG else rs[0] - 1
        if flag & capi.INDEL_FLAG:
            s, e = A + 1, B - 1
        elif flag & capi.TDUPLICATION_FLAG:
            s, e = B, A
        elif flag & capi.INVL_FLAG:
            s, e = A + 1, B
        else:
            s, e = A, B - 1
        L2, R2 = e2 - 1, rs[1] - 1
        out["homo_at1"] = run_at(seq1, s, e + 1); out["homo_at2"] = run_at(seq2, L2 + 1, R2)
        out["homo_out"] = [run_outside(seq1, s - 1, e + 1), run_outside(seq2, L2, R2)]
        out["homo_in"] = [run_inside(seq1, s, e), run_inside(seq2, L2 + 1, R2 - 1)]
    return out


def set_counts(r, c: dict):
    """fill capi.AgeResult.counts from a counts dict"""
    for k in range(2):
        r.counts.n_aligned[k], r.counts.n_identic[k], r.counts.n_gap[k] = c["n_aligned"][k], c["n_identic"][k], c["n_gap"][k]
    for k in range(3):
        r.counts.homo_at1[k], r.counts.homo_at2[k] = c["homo_at1"][k], c["homo_at2"][k]
    r.counts.homo_out1, r.counts.homo_out2 = c["homo_out"]
    r.counts.homo_in1, r.counts.homo_in2 = c["homo_in"]


def make_job(seq1: bytes, seq2: bytes, flag, match, mismatch, go, ge):
    j = capi.AgeJob()
    j.seq1, j.len1, j.seq2, j.len2 = seq1, len(seq1), seq2, len(seq2)
    j.flag, j.match, j.mismatch, j.gap_open, j.gap_extend = flag, match, mismatch, go, ge
    j.partner = -1
    return j


def make_view(seq: bytes, name: str, start: int, reverse: bool):
    v = capi.AgeSeqView()
    v.seq, v.len, v.name, v.start, v.reverse = seq, len(seq), name.encode("latin-1"), start, int(reverse)
    return v


def revcom_view(s: dict) -> dict:
    """Sequence::clone()+revcom() (Sequence.hh:134-136,190-198) on a fixture sequence dict."""
    n = len(s["seq"])
    start = s["start"] - (n - 1) if s["reverse"] else s["start"] + (n - 1)
    return {"name": s["name"], "start": start, "reverse": not s["reverse"], "seq": revcom_str(s["seq"])}


def candidate_record(case: dict, engine) -> str:
    """What age_align prints for one candidate (age_align.cpp:263-283 / Trash/age_align.cpp:322-343):
    `engine(seq1, seq2, flag, m, M, go, ge)` -> summary dict.  -both picks aligner1 iff score1 >= score2."""
    s1, s2 = case["s1"], case["s2"]
    sc = (case["match"], case["mismatch"], case["go"], case["ge"])
    variants = [s2] + ([revcom_view(s2)] if case["both"] else [])
    res = [engine(s1["seq"].encode(), v["seq"].encode(), case["flag"], *sc) for v in variants]
    if all(r["status"] != 0 for r in res):
        return ""
    pick = 0
    if case["both"] and not (res[0]["status"] == 0 and (res[1]["status"] != 0 or res[0]["score"] >= res[1]["score"])):
        pick = 1
    return format_record(s1, variants[pick], case, res[pick])


def format_record(s1: dict, s2: dict, case: dict, summary: dict) -> str:
    r, keep = summary_to_result(summary)
    b1, b2 = s1["seq"].encode(), s2["seq"].encode()
    job = make_job(b1, b2, case["flag"], case["match"], case["mismatch"], case["go"], case["ge"])
    return capi.format_alignment(make_view(b1, s1["name"], s1["start"], s1["reverse"]),
                                 make_view(b2, s2["name"], s2["start"], s2["reverse"]), job, r)


def load_cases():
    return json.loads((GOLDEN / "aligner_cases.json").read_text())["cases"]


def strip_time(text: str) -> str:
    return "".join(l for l in text.splitlines(keepends=True) if not l.startswith("Alignment time is"))
