"""Synthetic SV candidates for the benchmark configurations of BASELINE.json.

All generators are seeded (numpy default_rng), draw bases uniformly from ACGT, mutate the contig with 1 %
substitutions and 0.1 % one-base indels, and reverse-complement half of the contigs so that `-both` matters.
Each candidate is (window: bytes, contig: bytes).  The same candidates can be written as FASTA + the 7-column
variant table of the batch CLI (age_align.cpp:169-190) so that one set of files feeds the GPU path and the
reference CPU binary alike.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[:] = np.arange(256, dtype=np.uint8)
for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
    _COMP[a] = b


def revcom(a: np.ndarray) -> np.ndarray:
    return _COMP[a[::-1]]


def _mutate(rng, a: np.ndarray, sub=0.01, indel=0.001) -> np.ndarray:
    a = a.copy()
    n = a.size
    m = rng.random(n) < sub
    a[m] = _ACGT[rng.integers(0, 4, int(m.sum()))]
    dele = rng.random(n) < indel / 2
    ins = (rng.random(n) < indel / 2) & ~dele
    if dele.any() or ins.any():
        reps = (~dele).astype(np.int64) + ins            # deleted bases vanish, a random base follows the inserted-after ones
        out = np.repeat(a, reps)
        at = np.cumsum(reps)[ins] - 1
        out[at] = _ACGT[rng.integers(0, 4, at.size)]
        a = out
    return a


def _logu(rng, lo, hi) -> int:
    return int(np.exp(rng.uniform(np.log(lo), np.log(hi))))


def indel_candidates(n: int, seed: int = 1002, window: int = 5000, flank: int = 500, max_del: int = 4000, max_ins: int = 1000,
                     p_del: float = 0.5, min_sv: int = 50):
    """C2: 50 % deletions (contig = 2 flanks around the deleted segment), 50 % insertions."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        w = _ACGT[rng.integers(0, 4, window)]
        if rng.random() < p_del:
            L = min(_logu(rng, min_sv, max_del), window - 2 * flank)
            s = int(rng.integers(flank, window - flank - L + 1))
            q = np.concatenate([w[s - flank:s], w[s + L:s + L + flank]])
        else:
            p = int(rng.integers(flank, window - flank + 1))
            L = _logu(rng, min_sv, max_ins)
            q = np.concatenate([w[p - flank:p], _ACGT[rng.integers(0, 4, L)], w[p:p + flank]])
        q = _mutate(rng, q)
        if rng.random() < 0.5:
            q = revcom(q)
        out.append((w.tobytes(), q.tobytes()))
    return out


def tdup_candidates(n: int, seed: int = 1003, window: int = 4000, flank: int = 500):
    """C3 (tdup part): contig spans the junction of a tandem duplication of W[a:b]."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        w = _ACGT[rng.integers(0, 4, window)]
        L = int(rng.integers(flank + 50, window // 2))
        a = int(rng.integers(10, window - L - 10))
        b = a + L
        q = np.concatenate([w[b - flank:b], w[a:a + flank]])
        q = _mutate(rng, q)
        if rng.random() < 0.5:
            q = revcom(q)
        out.append((w.tobytes(), q.tobytes()))
    return out


C3_WINDOWS = (1000, 2000, 4000, 8000)
C3_CONTIGS = (500, 1000, 2000)


def mixed_candidates(n: int, seed: int = 1003, p_indel: float = 0.7):
    """C3: len1 in {1, 2, 4, 8} kb x len2 in {0.5, 1, 2} kb, uniform over the 12 bins; 70 % indel candidates as in C2,
    30 % tandem duplications.  Returns (kind, window, contig) with kind "indel" / "tdup": the reference runs one mode
    per invocation (age_align.cpp:207-221), so the two kinds go to two invocations (-i -b / -d -b)."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        window = C3_WINDOWS[int(rng.integers(0, 4))]
        len2 = C3_CONTIGS[int(rng.integers(0, 3))]
        sub_seed = int(rng.integers(0, 2**31 - 1))
        if rng.random() < p_indel:
            flank = min(len2 // 2, (window - 100) // 2)
            w, q = indel_candidates(1, seed=sub_seed, window=window, flank=flank, max_del=max(60, window - 2 * flank),
                                    max_ins=max(60, min(1000, len2)))[0]
            out.append(("indel", w, q))
        else:
            flank = min(len2 // 2, window // 2 - 120)
            w, q = tdup_candidates(1, seed=sub_seed, window=window, flank=flank)[0]
            out.append(("tdup", w, q))
    return out


def inversion_candidates(n: int, seed: int = 1004, window: int = 20000, flank: int = 2000):
    """C4: inversion of W[a:b]; the contig holds `flank` bases outside and `flank` inside one breakpoint."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        w = _ACGT[rng.integers(0, 4, window)]
        L = int(rng.integers(flank, max(flank + 1, min(12000, window - 2 * flank - 2))))
        a = int(rng.integers(flank, window - flank - L))
        b = a + L
        inv = revcom(w[a:b])
        if rng.random() < 0.5:      # left breakpoint: outside-left + start of the inverted segment
            q = np.concatenate([w[a - flank:a], inv[:flank]])
        else:                       # right breakpoint
            q = np.concatenate([inv[-flank:], w[b:b + flank]])
        q = _mutate(rng, q)
        out.append((w.tobytes(), q.tobytes()))
    return out


def long_sv_candidates(n: int, seed: int = 1005, window: int = 50000, flank: int = 5000):
    """C5: long deletions, 10 kb contigs against 50 kb windows."""
    return indel_candidates(n, seed=seed, window=window, flank=flank, max_del=35000, p_del=1.0, min_sv=5000)


def write_files(cands, outdir: Path, prefix: str = "c", ids=None):
    """FASTA pair + variant table for the batch CLI: one target record and one query record per candidate, named after
    `ids` (default 0..n-1) so that shards of one candidate list keep the names of the whole list."""
    outdir = Path(outdir)
    outdir.mkdir(parents=True, exist_ok=True)
    t, q, v = outdir / f"{prefix}.target.fa", outdir / f"{prefix}.query.fa", outdir / f"{prefix}.var.txt"
    with open(t, "wb") as ft, open(q, "wb") as fq, open(v, "w") as fv:
        for k, (w, c) in enumerate(cands):
            i = k if ids is None else ids[k]
            ft.write(b">t%d\n%s\n" % (i, w))
            fq.write(b">q%d\n%s\n" % (i, c))
            fv.write(f"t{i}\t1\t{len(w)}\tq{i}\t1\t{len(c)}\tcand{i}\n")
    return t, q, v
