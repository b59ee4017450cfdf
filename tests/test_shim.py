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
