"""`.age` format conformance: the records the product prints must stay parseable by the reference's own
consumer, src/PipTools/ASV_GetExcisedRegionFromAGE.pl:85-190, which walks the text with a fixed sequence of regular
expressions and positional `<I>` reads.  `consume()` below is that state machine with the same patterns (Perl -> re);
it is run over every golden case as formatted by the PRODUCT formatter (age_format_alignment) and the numbers it
extracts are checked against the structured result (breakpoints / blocks) the text was printed from."""
import re

import pytest

import ageutil as U

CASES = U.load_cases()

RE_MATCH = re.compile(r"^MATCH\s*=")
RE_NUCS = re.compile(r"nucs\s+'(\S+).*'$")
RE_IDENTIC = re.compile(r"^Identic:\s*(\d+)\s*\(\s*(\d+)%\)\s*nucs\s*=>\s*(\d+)\s*\(\s*(\d+)%\)\s*(\d+)\s*\(\s*(\d+)%\)")
RE_EXC = re.compile(r"\s+EXCISED\s+REGION\s+")
RE_EXC4 = re.compile(r"\s+\[\s*(\d+),\s*(\d+)\]\s+EXCISED\s+REGION\s+\[\s*(\d+),\s*(\d+)\]")
RE_IDENTITY = re.compile(r"^Identity\s+")
RE_CI = re.compile(r"\s+\[\s*(\d+),\s*(\d+)\]\s+to\s+\[\s*(\d+),\s*(\d+)\]")


class FormatError(AssertionError):
    pass


def consume(text: str) -> dict:
    """ASV_GetExcisedRegionFromAGE.pl:70-190 for one record; raises where the Perl script would die."""
    lines = text.split("\n")
    it = iter(lines)
    rec = {"flag": 0}
    for line in it:
        if RE_MATCH.match(line):                                         # :87-100
            if rec["flag"]:
                raise FormatError("second MATCH line")
            next(it)                                                     # "<I>; # Space line"
            first, second = next(it), next(it)
            m1, m2 = RE_NUCS.search(first), RE_NUCS.search(second)
            if not m1 or not m2:
                raise FormatError("First/Second seq lines do not end in nucs '<id>...'")
            rec["tarId"], rec["qryId"] = m1.group(1), m2.group(1)
            rec["flag"] = 1
        elif RE_IDENTIC.match(line):                                     # :101-108
            if not rec["flag"]:
                raise FormatError("Identic before MATCH")
            rec["identic"] = tuple(int(x) for x in RE_IDENTIC.match(line).groups())
        elif line.startswith("Alignment:"):                              # :109-
            if not rec["flag"]:
                raise FormatError("Alignment before MATCH")
            first, second = next(it), next(it)
            if RE_EXC.search(first) and RE_EXC.search(second):
                a, b = RE_EXC4.search(first), RE_EXC4.search(second)
                if not a or not b:
                    raise FormatError("EXCISED REGION coordinates")
                rec["t"] = tuple(int(x) for x in a.groups())
                rec["q"] = tuple(int(x) for x in b.groups())
                n, cur = 0, line
                while not RE_IDENTITY.match(cur):                        # :165 "I don't care these lines", at most 10
                    cur = next(it); n += 1
                    if n > 10:
                        raise FormatError("no 'Identity' line within 10 lines of the alignment ranges")
                ci = []
                for k in range(3):                                       # at / outside / inside breakpoints
                    first, second = next(it), next(it)
                    m = RE_CI.search(first)
                    if m:
                        ci.append(tuple(int(x) for x in m.groups()))
                    if k < 2:
                        cur = next(it)
                        if not RE_IDENTITY.match(cur):
                            raise FormatError("Cannot find the word 'Identity'")
                rec["ci"] = ci
    return rec


def printed(case):
    """(text, summary, s2 view) of the aligner the product prints for a golden case (oracle result, product formatter)."""
    s1, s2 = case["s1"], case["s2"]
    sc = (case["match"], case["mismatch"], case["go"], case["ge"])
    variants = [s2] + ([U.revcom_view(s2)] if case["both"] else [])
    res = [U.oracle_aligner(s1["seq"].encode(), v["seq"].encode(), case["flag"], *sc) for v in variants]
    pick = 0
    if case["both"] and not (res[0]["status"] == 0 and (res[1]["status"] != 0 or res[0]["score"] >= res[1]["score"])):
        pick = 1
    if res[pick]["status"] != 0:
        return "", res[pick], variants[pick]
    return U.format_record(s1, variants[pick], case, res[pick]), res[pick], variants[pick]


def coord(view, pos):
    return view["start"] - (pos - 1) if view["reverse"] else view["start"] + (pos - 1)


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_record_is_consumable(case):
    text, summary, s2 = printed(case)
    if not text:
        pytest.skip("no alignment made")
    assert text.rstrip("\n") == case["expected"].rstrip("\n")
    rec = consume(text)
    assert rec["flag"] == 1
    assert rec["tarId"] == case["s1"]["name"].split()[0] and rec["qryId"] == s2["name"].split()[0]
    left, right = summary["left"], summary["right"]
    if left and right:                                   # two fragments: the consumer needs all of these
        assert "identic" in rec and "t" in rec and "q" in rec and "ci" in rec
        ave_base, ave_iden, lb, li, rb, ri = rec["identic"]
        assert lb + rb == ave_base and 0 <= ave_iden <= 100 and 0 <= li <= 100 and 0 <= ri <= 100
        if case["flag"] == 0x01:                         # indel: both fragments on seq1 itself
            ls1, ls2, _ = left[0]
            le1, le2 = left[-1][0] + left[-1][2] - 1, left[-1][1] + left[-1][2] - 1
            rs1, rs2, _ = right[0]
            re1, re2 = right[-1][0] + right[-1][2] - 1, right[-1][1] + right[-1][2] - 1
            assert rec["t"] == (coord(case["s1"], ls1), coord(case["s1"], le1), coord(case["s1"], rs1), coord(case["s1"], re1))
            assert rec["q"] == (coord(s2, ls2), coord(s2, le2), coord(s2, rs2), coord(s2, re2))
            # the first/last block ends are the breakpoints (bp[0] = lg1, lg2, rg1, rg2)
            assert (le1, le2, rs1, rs2) == tuple(summary["bp"][0])
    else:
        assert "t" not in rec                            # single fragment: no EXCISED REGION on the range lines


def test_consumer_rejects_drift():
    """the state machine is not vacuous: typical drifts of the printer are caught"""
    case = next(c for c in CASES if c["id"] == "Test2_0")
    text, _, _ = printed(case)
    consume(text)
    with pytest.raises((FormatError, StopIteration)):
        consume(text.replace("Identity outside breakpoints", "Outside identity"))
    with pytest.raises(FormatError):
        consume(text.replace("' \n", "\n").replace("nucs '", "nucs "))
    assert "t" not in consume(text.replace(" EXCISED REGION ", " EXCISED  "))
