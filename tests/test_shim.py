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
