"""The C++ in-process interface (include/age_shim.hh: Sequence / Scorer / AGEaligner shaped like the reference's classes)
through a program written like the reference's candidate loop (examples/shim_example.cpp)."""
import subprocess
from pathlib import Path

import pytest

import ageutil as U
from asmvar_b200 import capi

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_shim_example_matches_reference_record(tmp_path):
    exe = tmp_path / "shim_example"
    subprocess.run(["g++", "-std=c++17", "-I", str(ROOT / "include"), str(ROOT / "examples" / "shim_example.cpp"),
                    f"-L{ROOT / 'asmvar_b200'}", "-lage_b200", f"-Wl,-rpath,{ROOT / 'asmvar_b200'}", "-o", str(exe)], check=True)
    src = next(c for c in U.load_cases() if c["id"] == "Test2_3")
    for s2 in (src["s2"]["seq"], U.revcom_str(src["s2"]["seq"])):          # both strands: exercises the -both choice
        case = {"s1": {"name": "first", "start": 1, "reverse": False, "seq": src["s1"]["seq"]},
                "s2": {"name": "second", "start": 1, "reverse": False, "seq": s2},
                "flag": capi.INDEL_FLAG, "both": True, "match": 1, "mismatch": -1, "go": -10, "ge": -1}
        want = U.candidate_record(case, U.oracle_aligner)
        got = subprocess.run([str(exe), case["s1"]["seq"], case["s2"]["seq"]], check=True, capture_output=True, text=True).stdout
        assert got == want


def test_shim_align_result_matches_reference_dumps(tmp_path):
    """SetAlignResult()/align_result() of include/age_shim.hh on the GPU for every golden AlignResult case
    (examples/shim_align_result.cpp prints the same JSON as the dump driver of the reference aligner)."""
    import json
    import test_align_result as T
    exe = tmp_path / "shim_align_result"
    subprocess.run(["g++", "-std=c++17", "-I", str(ROOT / "include"), str(ROOT / "examples" / "shim_align_result.cpp"),
                    f"-L{ROOT / 'asmvar_b200'}", "-lage_b200", f"-Wl,-rpath,{ROOT / 'asmvar_b200'}", "-o", str(exe)], check=True)
    lines = []
    for g in T.GOLD:
        case, s1, s2 = T.inputs(g)
        f = [g["id"], s1["name"], s1["start"], int(s1["reverse"]), s1["seq"], s2["name"], s2["start"], int(s2["reverse"]), s2["seq"],
             case["flag"], case["match"], case["mismatch"], case["go"], case["ge"]]
        lines.append("\t".join(str(x) for x in f))
    p = subprocess.run([str(exe)], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True)
    got = [json.loads(l) for l in p.stdout.splitlines() if l.strip()]
    assert len(got) == len(T.GOLD)
    for a, g in zip(got, T.GOLD):
        assert a["id"] == g["id"] and a["aligned"] == g["aligned"]
        if not g["aligned"]:
            continue
        assert bool(a["printed_bytes"]) == bool(g["printed_bytes"])
        T.compare(a, g)
        assert a["id1"] == g["id1"] and a["id2"] == g["id2"]


def _recall_cases():
    import json
    import numpy as np
    d = json.loads((U.GOLDEN / "recall_cases.json").read_text())
    for c in d["cases"]:
        if "regen" in c:                                   # the window beyond the 10 GB guard: any sequence of that size
            rng = np.random.default_rng(c["regen"]["seed"])
            c["tseq"] = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=c["regen"]["tlen"]).tobytes()).decode()
            c["qseq"] = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=c["regen"]["qlen"]).tobytes()).decode()
    return d["keys"], d["cases"]


def test_shim_realign_and_recall_matches_reference(tmp_path):
    """VarUnit::ReAlignAndReCallVar of include/age_varunit.hh (window extension, 10 GB guard, -both choice, the
    `# type:strand:...` header + record, variant re-calling from the excised region) against the outputs of the
    unmodified AsmvarDetect sources (tests/golden/recall_cases.json); one candidate per call, then batched."""
    import json
    exe = tmp_path / "shim_recall"
    subprocess.run(["g++", "-std=c++17", "-I", str(ROOT / "include"), str(ROOT / "examples" / "shim_recall.cpp"),
                    f"-L{ROOT / 'asmvar_b200'}", "-lage_b200", f"-Wl,-rpath,{ROOT / 'asmvar_b200'}", "-o", str(exe)], check=True)
    keys, cases = _recall_cases()
    line = lambda c: "\t".join(str(c[k]) for k in keys)

    def check_vus(got, want, cid):
        assert len(got) == len(want), cid
        for a, b in zip(got, want):
            for k in b:
                if k == "mismap":
                    assert abs(a[k] - b[k]) <= 1e-12 * max(1.0, abs(b[k])), (cid, k)     # printed with 6 significant digits on both sides
                else:
                    assert a[k] == b[k], (cid, k, a[k], b[k])

    # one candidate per call
    p = subprocess.run([str(exe)], input="\n".join(line(c) for c in cases) + "\n", capture_output=True, text=True, check=True)
    got = [json.loads(l) for l in p.stdout.splitlines() if l.strip()]
    assert len(got) == len(cases)
    for g, c in zip(got, cases):
        assert g["id"] == c["id"]
        if c["expected"] is None:
            # the reference dereferences an empty _map here (undefined behaviour, it crashes): we print and re-call nothing
            assert g["success"] == 1 and g["vus"] == [] and g["text"].count("\nMATCH = ") == 2
            continue
        assert g["success"] == c["expected"]["success"], c["id"]
        assert g["text"] == c["expected"]["text"], c["id"]
        check_vus(g["vus"], c["expected"]["vus"], c["id"])
    # the whole loop as GPU batches (consecutive cases with equal options share a batch)
    pinned = [c for c in cases if c["expected"] is not None]
    p = subprocess.run([str(exe), "batch"], input="\n".join(line(c) for c in pinned) + "\n", capture_output=True, text=True, check=True)
    out = [json.loads(l) for l in p.stdout.splitlines() if l.strip()]
    assert len(out) == len(pinned) + 1
    for g, c in zip(out, pinned):
        assert g["id"] == c["id"] and g["success"] == c["expected"]["success"]
        check_vus(g["vus"], c["expected"]["vus"], c["id"])
    assert out[-1]["all_text"] == "".join(c["expected"]["text"] for c in pinned)
