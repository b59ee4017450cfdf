"""The three "identity at / outside / inside breakpoints" numbers of printAlignment and SetAlignResult
(AGEaligner.cpp:392-447), which the product computes in one pass (prefix function) instead of the reference's
restart per candidate length.  Checked here against the plain definition -- the longest stretch before the left
breakpoint that is repeated after the right one, and so on -- on repeat-rich sequences with N and lower case, where
Sequence::sameNuc is not an equivalence (N matches nothing, not even N).  CPU only."""
import numpy as np

import ageutil as U
from asmvar_b200 import capi


def same(a, b):
    a, b = a.upper(), b.upper()
    return a == b and a in "ACGTU"


def ch(s, i):
    return s[i] if 0 <= i < len(s) else "\0"


def outside(s, bp1, bp2):
    if bp1 >= bp2:
        return 0
    for m in range(min(bp1 + 1, len(s) - bp2), 0, -1):
        if all(same(ch(s, bp1 - m + 1 + k), ch(s, bp2 + k)) for k in range(m)):
            return m
    return 0


def inside(s, bp1, bp2):
    if bp1 >= bp2:
        return 0
    for m in range((bp2 - bp1 + 1) // 2, 0, -1):
        if all(same(ch(s, bp1 + k), ch(s, bp2 - m + 1 + k)) for k in range(m)):
            return m
    return 0


def at(s, bp1, bp2):
    if bp1 >= bp2:
        return 0
    d = bp2 - bp1
    down = up = 0
    i = bp1 - 1
    while i >= 0 and same(ch(s, i), ch(s, i + d)):
        down += 1; i -= 1
    i = bp1
    while i < len(s) and same(ch(s, i), ch(s, i + d)):
        up += 1; i += 1
    return up + down


def test_one_pass_identities_match_the_definition():
    rng = np.random.default_rng(2024)
    n_two = 0
    for it in range(400):
        unit = "".join(rng.choice(list("ACGTNacn"), size=int(rng.integers(1, 7))))
        s1 = (unit * 60)[: int(rng.integers(40, 160))]
        s1 = "".join(c if rng.random() > 0.04 else str(rng.choice(list("ACGTN"))) for c in s1)
        a = int(rng.integers(3, len(s1) - 12))                 # left fragment = s1[0:a]
        c = int(rng.integers(a + 2, len(s1) - 3))              # right fragment starts at 1-based c+1
        b = int(rng.integers(2, len(s1) - c + 1))
        s2 = s1[:a] + s1[c:c + b]
        summary = {"status": 0, "score": a + b, "flag": capi.INDEL_FLAG, "used_aux": 0, "main_score": a + b, "aux_score": -1,
                   "bp": [(a, a, c + 1, a + 1)], "alt_index": -1, "left": [(1, 1, a)], "right": [(c + 1, a + 1, b)], "alt_left": [], "alt_right": []}
        r, keep = U.summary_to_result(summary)
        got = capi.describe_alignment(U.make_view(s1.encode(), "t", 1, False), U.make_view(s2.encode(), "q", 1, False), r)
        if len(got["map"]) < 2:
            continue
        n_two += 1
        # first sequence: fragments end at a and restart at c + 1 (1-based) -> 0-based breakpoints as printAlignment derives them
        want1 = [at(s1, a, c), inside(s1, a, c - 1), outside(s1, a - 1, c)]
        # second sequence: the fragments abut (end a, start a + 1)
        want2 = [at(s2, a, a), inside(s2, a, a - 1), outside(s2, a - 1, a)]
        assert got["homo_run"] == [want1[0], want2[0], want1[1], want2[1], want1[2], want2[2]], (it, s1, a, c, b)
    assert n_two > 300


def identity_lines(text: str):
    """the six numbers of the three 'Identity ... breakpoints' paragraphs of a record: [at1, at2, out1, out2, in1, in2]"""
    lines = text.split("\n")
    out = []
    for title in ("Identity at breakpoints: ", "Identity outside breakpoints: ", "Identity inside breakpoints: "):
        i = lines.index(title)
        for k in (1, 2):
            out.append(int(lines[i + k].split("=>")[1].split("nucs")[0]))
    return out


def test_formatter_identities_on_repeats_match_the_definition():
    """printAlignment's path (age_format_alignment) tries the candidate lengths longest-first on nucleotide classes with a
    budget of comparisons and falls back to the prefix function: repeat-rich and low-complexity sequences of up to 3 kb (where
    many lengths fail late and the budget runs out) against the plain definition."""
    rng = np.random.default_rng(77)
    n_two = n_long = 0
    for it in range(160):
        kind = it % 4
        n = int(rng.integers(60, 400)) if kind < 2 else int(rng.integers(800, 3000))
        if kind in (0, 2):      # periodic with noise
            unit = "".join(rng.choice(list("ACGTNacn"), size=int(rng.integers(1, 9))))
            s1 = (unit * (n // len(unit) + 1))[:n]
            s1 = "".join(c if rng.random() > 0.01 else str(rng.choice(list("ACGT"))) for c in s1)
        else:                   # homopolymer with a few other bases: nearly every length matches for a long while
            s1 = "".join("A" if rng.random() > 0.004 else str(rng.choice(list("CGTn"))) for _ in range(n))
        a = int(rng.integers(3, len(s1) - 12))
        c = int(rng.integers(a + 2, len(s1) - 3))
        b = int(rng.integers(2, len(s1) - c + 1))
        s2 = s1[:a] + s1[c:c + b]
        summary = {"status": 0, "score": a + b, "flag": capi.INDEL_FLAG, "used_aux": 0, "main_score": a + b, "aux_score": -1,
                   "bp": [(a, a, c + 1, a + 1)], "alt_index": -1, "left": [(1, 1, a)], "right": [(c + 1, a + 1, b)], "alt_left": [], "alt_right": []}
        r, keep = U.summary_to_result(summary)
        v1, v2 = U.make_view(s1.encode(), "t", 1, False), U.make_view(s2.encode(), "q", 1, False)
        text = capi.format_alignment(v1, v2, U.make_job(s1.encode(), s2.encode(), capi.INDEL_FLAG, 1, -1, -10, -1), r)
        if "Identity at breakpoints" not in text:
            continue
        n_two += 1
        n_long += n >= 800
        want = [at(s1, a, c), at(s2, a, a), outside(s1, a - 1, c), outside(s2, a - 1, a), inside(s1, a, c - 1), inside(s2, a, a - 1)]
        assert identity_lines(text) == want, (it, kind, n, a, c, b)
    assert n_two > 120 and n_long > 50
