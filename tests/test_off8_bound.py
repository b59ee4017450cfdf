"""The invariant behind the 8-bit storage of the forward-maxima plane ("D as 8-bit row offsets"):
with the reference recurrence (AGEaligner.cpp:988-1028 fill, :873-917 prefix maxima) consecutive rows of Fmax differ by
at most smax = max(match, mismatch, 0), in both directions of the matrix.  Checked here on a direct restatement of the
recurrence in numpy-free Python (small matrices, many scorings), independently of the CUDA code and of the oracle."""
import random

import pytest

H, V, D, STOP = 2, 3, 1, 0


def fill_fmax(x, y, match, mismatch, go, ge):
    """Fmax[r][c] for seq1 = x (columns), seq2 = y (rows), borders zero."""
    n1, n2 = len(x), len(y)
    F = [[0] * (n1 + 1) for _ in range(n2 + 1)]
    T = [[STOP] * (n1 + 1) for _ in range(n2 + 1)]
    acgt = set("ACGTacgtUu")

    def S(a, b):
        if a not in acgt or b not in acgt:
            return 0
        return match if a.upper() == b.upper() and a.upper() in "ACGTU" else mismatch

    for r in range(1, n2 + 1):
        for c in range(1, n1 + 1):
            d = F[r - 1][c - 1] + S(y[r - 1], x[c - 1])
            h = F[r][c - 1] + (ge if T[r][c - 1] in (H, V) else go)
            v = F[r - 1][c] + (ge if T[r - 1][c] in (H, V) else go)
            s, t = d, D
            if h >= s:
                s, t = h, H
            if v >= s:
                s, t = v, V
            if s < 0:
                s, t = 0, STOP
            F[r][c], T[r][c] = s, t
    M = [[0] * (n1 + 1) for _ in range(n2 + 1)]
    for r in range(1, n2 + 1):
        for c in range(1, n1 + 1):
            M[r][c] = max(F[r][c], M[r - 1][c], M[r][c - 1])
    return M


@pytest.mark.parametrize("scoring", [(1, -1, -10, -1), (1, -2, -2, -1), (2, -3, -5, -2), (15, -4, -3, -3), (2, 9, -3, -2),
                                     (7, -2, -1, -1), (1, -1, -1, -3), (3, 3, -1, -1)])
def test_row_and_column_steps_of_fmax_are_bounded(scoring):
    match, mismatch, go, ge = scoring
    smax = max(match, mismatch, 0)
    rng = random.Random(hash(scoring) & 0xFFFF)
    for trial in range(6):
        alphabet = "ACGT" if trial % 2 == 0 else "ACGTNacgu"
        n1, n2 = rng.randint(20, 70), rng.randint(20, 70)
        x = "".join(rng.choice(alphabet) for _ in range(n1))
        # a contig that shares long stretches with the window, so that the maxima really grow
        y = "".join(x[i] if i < n1 and rng.random() < 0.9 else rng.choice(alphabet) for i in range(n2))
        M = fill_fmax(x, y, match, mismatch, go, ge)
        worst = 0
        for r in range(1, n2 + 1):
            for c in range(1, n1 + 1):
                assert 0 <= M[r][c] - M[r - 1][c] <= smax, (scoring, r, c)
                assert 0 <= M[r][c] - M[r][c - 1] <= smax, (scoring, r, c)
                worst = max(worst, M[r][c] - M[r - 1][c])
        # 8 rows of a lane: offsets from the row above the lane's first row stay below 256 for smax <= 15 (in 2 * score units)
        assert 2 * smax * 7 <= 255 or smax > 15
