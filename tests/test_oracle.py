"""Pins the CPU oracle (oracle/age_oracle.c) and the product's `.age` formatter to the reference:
every golden record (the reference's own 15 known-answer cases, example/Test1, example/Test2 and
220 randomised cases produced by the compiled reference binary) must be reproduced byte-for-byte
by  oracle result -> age_format_alignment."""
import pytest

import ageutil as U

CASES = U.load_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_oracle_reproduces_reference_record(case):
    got = U.candidate_record(case, U.oracle_aligner)
    assert got.rstrip("\n") == case["expected"].rstrip("\n")


def test_case_inventory():
    ids = {c["id"] for c in CASES}
    assert {f"example{i}" for i in range(1, 16)} <= ids          # examples/TestDecription_v0.4.txt
    assert "Test1_orig" in ids and {f"Test2_{i}" for i in range(11)} <= ids
    modes = {c["flag"] for c in CASES}
    assert modes == {0x01, 0x02, 0x04, 0x08, 0x10}
    assert any(c["both"] for c in CASES) and any(c["s1"]["reverse"] for c in CASES) and any(c["s2"]["reverse"] for c in CASES)
