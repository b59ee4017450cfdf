#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/recall_cases.json: inputs and outputs of the unmodified
VarUnit::ReAlignAndReCallVar (src/AsmvarDetect/VarUnit.cpp:66-82, via oracle/_ref/age_recall_dump built by
oracle/build_ref.sh) on seeded synthetic variants: deletions, insertions, replacements, SNPs / MNPs and clean windows,
on both strands, several flank extensions (clipped at the record ends), all five modes, and one window beyond the
reference's 10 GB guard.  Run in the build container only."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
COMP = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")


def rc(s: bytes) -> bytes:
    return s.translate(COMP)[::-1]


def rand(rng, n):
    return bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n).tobytes())


def noisy(rng, s: bytes, p=0.004) -> bytes:
    a = bytearray(s)
    for i in np.nonzero(rng.random(len(a)) < p)[0]:
        a[i] = b"ACGT"[(b"ACGT".index(bytes([a[i]]).upper()) + 1 + int(rng.integers(0, 3))) % 4]
    return bytes(a)


def main():
    rng = np.random.default_rng(20260)
    cases = []
    kinds = ["del", "ins", "repl", "snp", "mnp", "none", "del", "ins"]
    scorings = [(1, -1, -10, -1), (1, -1, -10, -1), (2, -3, -5, -2), (1, -2, -2, -1)]
    for i in range(84):
        n = int(rng.integers(600, 2000))
        t = rand(rng, n)
        pos = int(rng.integers(200, n - 200))
        kind = kinds[i % len(kinds)]
        L = int(rng.integers(1, 160))
        if kind == "del":
            q, tlen, qlen = t[:pos] + t[pos + L:], L, 0
        elif kind == "ins":
            q, tlen, qlen = t[:pos] + rand(rng, L) + t[pos:], 0, L
        elif kind == "repl":
            L2 = int(rng.integers(1, 120))
            q, tlen, qlen = t[:pos] + rand(rng, L2) + t[pos + L:], L, L2
        elif kind == "snp":
            q, tlen, qlen = t[:pos] + rc(t[pos:pos + 1]) + t[pos + 1:], 1, 1
        elif kind == "mnp":
            L = int(rng.integers(2, 9))
            q, tlen, qlen = t[:pos] + rc(t[pos:pos + L])[::-1] + t[pos + L:], L, L
        else:
            q, tlen, qlen = t, 0, 0
        if i % 5 == 3:
            q = noisy(rng, q)
        qpos = pos
        strand = "+"
        if i % 3 == 1:                                    # contig on the other strand: -both has to find it
            q = rc(q); strand = "-"
            qpos = len(q) - pos - qlen
        jit = int(rng.integers(0, 12))
        tstart, tend = max(1, pos - jit), min(n, pos + tlen + jit)
        qstart, qend = max(1, qpos - jit), min(len(q), qpos + qlen + jit)
        flag = 0x01 if i % 7 else [0x02, 0x04, 0x10, 0x08][(i // 7) % 4]
        both = 0 if i % 11 == 5 else 1
        ext = [50, 200, 500, 120][i % 4]
        m, mm, go, ge = scorings[i % len(scorings)]
        cases.append({"id": f"recall{i}", "tid": f"chr{i % 4 + 1}", "tseq": t.decode(), "qid": f"scaffold{i}", "qseq": q.decode(),
                      "tstart": tstart, "tend": tend, "qstart": qstart, "qend": qend, "type": kind.upper(), "strand": strand,
                      "score": int(rng.integers(50, 3000)), "mismap": [1e-10, 0.5, 0.003, 1.0][i % 4], "extend": ext, "both": both,
                      "flag": flag, "match": m, "mismatch": mm, "go": go, "ge": ge})
    # beyond the 10 GB guard of VarUnit.cpp:605-609 (5*n*m/1e9 + 0.5 > 10): must be refused without aligning
    t, q = rand(rng, 50000), rand(rng, 40000)
    cases.append({"id": "recall_huge", "tid": "chrH", "tseq": t.decode(), "qid": "scaffoldH", "qseq": q.decode(), "tstart": 1, "tend": 50000,
                  "qstart": 1, "qend": 40000, "type": "DEL", "strand": "+", "score": 100, "mismap": 0.001, "extend": 500, "both": 1,
                  "flag": 0x01, "match": 1, "mismatch": -1, "go": -10, "ge": -1})
    keys = ["id", "tid", "tseq", "qid", "qseq", "tstart", "tend", "qstart", "qend", "type", "strand", "score", "mismap", "extend", "both",
            "flag", "match", "mismatch", "go", "ge"]
    exe = str(ROOT / "oracle" / "_ref" / "age_recall_dump")
    for c in cases:                                       # one process per case: the reference can crash (see below)
        line = "\t".join(str(c[k]) for k in keys) + "\n"
        p = subprocess.run([exe], input=line, capture_output=True, text=True)
        if p.returncode != 0:
            # -inv when the auxiliary INVR aligner wins: SetAlignResult() leaves AlignResult empty and VarReCall() reads
            # _map[0] of an empty vector (VarUnit.cpp:361) -- undefined behaviour, SIGSEGV with g++ -O2
            assert c["flag"] == 0x04, (c["id"], p.returncode, p.stderr[-500:])
            c["expected"] = None
            c["reference_crashes"] = p.returncode
            continue
        o = json.loads(p.stdout.strip().splitlines()[-1])
        assert c["id"] == o["id"]
        c["expected"] = {"success": o["success"], "text": o["text"], "vus": o["vus"]}
    cases[-1]["tseq"] = cases[-1]["qseq"] = ""           # the refused window is regenerated by the test (keeps the fixture small)
    cases[-1]["regen"] = {"seed": 77, "tlen": 50000, "qlen": 40000}
    dst = ROOT / "tests" / "golden" / "recall_cases.json"
    dst.write_text(json.dumps({"generator": "oracle/make_golden_recall.py (oracle/_ref/age_recall_dump = src/AsmvarDetect/VarUnit.cpp, unmodified)",
                               "keys": keys, "cases": cases}, separators=(",", ":")))
    import collections
    ok = [c for c in cases if c["expected"]]
    print(collections.Counter(v["type"] for c in ok for v in c["expected"]["vus"]), sum(c["expected"]["success"] for c in ok), "aligned of", len(cases),
          "; reference crashed on", [c["id"] for c in cases if not c["expected"]])
    print(f"-> {dst} ({dst.stat().st_size} bytes)")


if __name__ == "__main__":
    main()
