#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY -- generates the golden fixtures under tests/golden/.

Run in the build container (where /root/reference exists) after oracle/build_ref.sh:

    python oracle/make_golden.py

Everything it writes is derived from (a) the reference's own golden files
(src/AsmvarAGE/examples/TestDecription_v0.4.txt, src/AsmvarAGE/src/example/Test{1,2}) and
(b) outputs of the reference binaries compiled from /root/reference (oracle/_ref/), run here.
The GPU box has no /root/reference; tests there read only the committed fixtures.

Fixtures
  tests/golden/aligner_cases.json   per-aligner cases: the two Sequence objects exactly as the
                                    aligner sees them + options + the reference's record text
  tests/golden/cli/*                CLI-level cases: input files, command line, expected stdout
"""
from __future__ import annotations

import gzip
import json
import os
import re
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("AGE_REFERENCE_ROOT", "/root/reference"))
AGE = REF / "src" / "AsmvarAGE"
BIN_BATCH = ROOT / "oracle" / "_ref" / "age_align_ref"
BIN_ORIG = ROOT / "oracle" / "_ref" / "age_align_ref_orig"
OUT = ROOT / "tests" / "golden"

FLAGS = {"indel": 0x01, "tdup": 0x02, "inv": 0x04, "invr": 0x08, "invl": 0x10}
COMP = {"a": "t", "c": "g", "t": "a", "g": "c", "A": "T", "C": "G", "T": "A", "G": "C", "u": "a", "U": "A"}


def strip_time(text: str) -> str:
    return "".join(l for l in text.splitlines(keepends=True) if not l.startswith("Alignment time is"))


def revcom(s: str) -> str:
    return "".join(COMP.get(c, c) for c in reversed(s))


class Seq:
    """Sequence.hh semantics needed to describe what the aligner sees."""

    def __init__(self, seq, name, start=1, reverse=False):
        self.seq, self.name, self.start, self.reverse = seq, name, start, reverse

    def substr(self, start, end):  # Sequence.hh:200-215
        n = len(self.seq)
        if start <= 0:
            start = 1
        if end <= 0 or end > n:
            end = n
        s, e = start, end
        if self.reverse:
            s, e = n - end + 1, n - start + 1
        return Seq(self.seq[s - 1:e], self.name, self.start + start - 1, self.reverse)

    def revcom(self):  # Sequence.hh:190-198
        n = len(self.seq)
        self.start = self.start - (n - 1) if self.reverse else self.start + (n - 1)
        self.reverse = not self.reverse
        self.seq = revcom(self.seq)
        return self

    def view(self):
        return {"name": self.name, "start": self.start, "reverse": bool(self.reverse), "seq": self.seq}


def parse_orig_fasta(path):  # Sequence::parseSequences (Sequence.hh:287-342): name = whole header line
    seqs, name, buf = [], None, []
    for line in open(path).read().split("\n"):
        if line.startswith(">"):
            if name:
                seqs.append(Seq("".join(buf), name))
            name, buf = line[1:], []
        else:
            buf.append("".join(c for c in line if c not in " -"))
    if name:
        seqs.append(Seq("".join(buf), name))
    return seqs


def load_fa(path):  # Fa::Load (Backup/Fa.h:27-50): id = first token after '>'
    fa, cur = {}, None
    op = gzip.open if str(path).endswith(".gz") else open
    for line in op(path, "rt"):
        tok = line.split()
        if not tok:
            continue
        if tok[0][0] == ">":
            cur = tok[0][1:]
            fa[cur] = []
        else:
            fa[cur].append(tok[0])
    return {k: "".join(v) for k, v in fa.items()}


def run(cmd, cwd=None):
    p = subprocess.run([str(c) for c in cmd], cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    return p.stdout


def orig_case(cid, f1, f2, opts, cwd):
    """Run the original CLI on single-record FASTA files; describe the aligner inputs."""
    out = strip_time(run([BIN_ORIG] + opts + [f1, f2], cwd=cwd))
    s1 = parse_orig_fasta(Path(cwd) / f1)[0]
    s2 = parse_orig_fasta(Path(cwd) / f2)[0]
    match, mismatch, go, ge = 1, -2, -2, -1
    flag, both, c1, c2, r1, r2 = 0, False, (-1, -1), (-1, -1), False, False
    for o in opts:
        if o.startswith("-match="): match = int(o[7:])
        elif o.startswith("-mismatch="): mismatch = int(o[10:])
        elif o.startswith("-go="): go = int(o[4:])
        elif o.startswith("-ge="): ge = int(o[4:])
        elif o.startswith("-coor1="): a, b = o[7:].split("-"); c1 = (int(a) if a else -1, int(b) if b else -1)
        elif o.startswith("-coor2="): a, b = o[7:].split("-"); c2 = (int(a) if a else -1, int(b) if b else -1)
        elif o == "-both": both = True
        elif o == "-revcom1": r1 = True
        elif o == "-revcom2": r2 = True
        else: flag |= FLAGS[o[1:]]
    if flag == 0:
        flag = FLAGS["indel"]
    s1 = s1.substr(*c1); s2 = s2.substr(*c2)
    if r1: s1.revcom()
    if r2: s2.revcom()
    # the record = stdout minus the trailing blank line that follows the (stripped) time line
    return {"id": cid, "cli": "orig", "opts": opts, "flag": flag, "both": both,
            "match": match, "mismatch": mismatch, "go": go, "ge": ge,
            "s1": s1.view(), "s2": s2.view(), "expected": out}


# ---------------------------------------------------------------------------------------------
def reference_examples(cases):
    ex = AGE / "examples"
    desc = (ex / "TestDecription_v0.4.txt").read_text()
    blocks = re.split(r"^\* Test #(\d+)\n", desc, flags=re.M)
    for k in range(1, len(blocks), 2):
        num, body = int(blocks[k]), blocks[k + 1]
        m = re.search(r"Run as: \./age_align test\d+_\*\.fa(.*)\n", body)
        opts = m.group(1).split()
        golden = body[m.end():]
        case = orig_case(f"example{num}", f"test{num}_first.fa", f"test{num}_second.fa", opts, ex)
        # pin the compiled oracle to the reference's own golden text
        g = strip_time(golden).rstrip("\n")
        assert case["expected"].rstrip("\n") == g, f"reference binary != shipped golden for test {num}"
        cases.append(case)
    print(f"  15 TestDecription cases reproduce the shipped golden text")


def test1_example(cases):
    t1 = AGE / "src" / "example" / "Test1"
    case = orig_case("Test1_orig", "target.1.fa", "query.1.fa",
                     ["-indel", "-both", "-match=1", "-mismatch=-1", "-go=-10", "-ge=-1", "-coor1=1-1000", "-coor2=1-534"], t1)
    assert case["expected"].rstrip("\n") == strip_time((t1 / "t1.age").read_text()).rstrip("\n"), "Test1 t1.age mismatch"
    cases.append(case)
    print("  Test1/t1.age reproduced")


def split_batch_records(text):
    recs, cur = [], None
    for line in strip_time(text).splitlines(keepends=True):
        if line.startswith("# "):
            if cur is not None:
                recs.append(cur)
            cur = [line]
        elif cur is not None:
            cur.append(line)
    if cur is not None:
        recs.append(cur)
    return ["".join(r) for r in recs]


def test2_example(cases):
    t2 = AGE / "src" / "example" / "Test2"
    args = ["--indel", "--both", "--match=1", "--mismatch=-1", "--go=-10", "--ge=-1"]
    out = run([BIN_BATCH] + args + ["--target", t2 / "target.fa", "--query", t2 / "query.fa", t2 / "var.txt"])
    assert strip_time(out) == strip_time((t2 / "t.age").read_text()), "Test2 t.age mismatch"
    print("  Test2/t.age reproduced")
    tfa, qfa = load_fa(t2 / "target.fa"), load_fa(t2 / "query.fa")
    rows = [l.split() for l in (t2 / "var.txt").read_text().splitlines() if l.strip() and not l.startswith("#")]
    recs = split_batch_records(out)
    assert len(recs) == len(rows)
    for i, (row, rec) in enumerate(zip(rows, recs)):
        s1 = Seq(tfa[row[0]], row[0]).substr(int(row[1]), int(row[2]))
        s2 = Seq(qfa[row[3]], row[3]).substr(int(row[4]), int(row[5]))
        body = rec.split("\n", 1)[1]
        cases.append({"id": f"Test2_{i}", "cli": "batch", "opts": args, "flag": 1, "both": True,
                      "match": 1, "mismatch": -1, "go": -10, "ge": -1, "s1": s1.view(), "s2": s2.view(),
                      "expected": body})
    # CLI-level fixture: same coordinates/names, everything outside the used windows masked with N (gz-friendly)
    cli = OUT / "cli" / "test2"
    cli.mkdir(parents=True, exist_ok=True)

    def masked(fa, used):
        out = {}
        for name, wins in used.items():
            s = fa[name]
            m = bytearray(b"N" * len(s))
            for a, b in wins:
                m[a - 1:b] = s[a - 1:b].encode()
            out[name] = m.decode()
        return out

    used_t, used_q = {}, {}
    for row in rows:
        used_t.setdefault(row[0], []).append((int(row[1]), int(row[2])))
        used_q.setdefault(row[3], []).append((int(row[4]), int(row[5])))
    for fn, fa in (("target.fa.gz", masked(tfa, used_t)), ("query.fa.gz", masked(qfa, used_q))):
        with gzip.GzipFile(cli / fn, "wb", mtime=0) as f:
            for name, s in fa.items():
                f.write(f">{name}\n".encode())
                for k in range(0, len(s), 60):
                    f.write(s[k:k + 60].encode() + b"\n")
    (cli / "var.txt").write_text((t2 / "var.txt").read_text())
    (cli / "cmd.txt").write_text(" ".join(args) + " --target target.fa.gz --query query.fa.gz var.txt\n")
    exp = run([BIN_BATCH] + args + ["--target", "target.fa.gz", "--query", "query.fa.gz", "var.txt"], cwd=cli)
    assert strip_time(exp) == strip_time(out), "masked Test2 fixture changes the output"
    (cli / "expected.age").write_text(strip_time(exp))
    print("  tests/golden/cli/test2 written (masked inputs give the identical t.age)")


# ---------------------------------------------------------------------------------------------
def mutate(rng, s, sub=0.02, indel=0.005):
    out = []
    for ch in s:
        u = rng.random()
        if u < sub:
            out.append("ACGT"[rng.integers(4)])
        elif u < sub + indel:
            continue
        elif u < sub + 2 * indel:
            out.append(ch); out.append("ACGT"[rng.integers(4)])
        else:
            out.append(ch)
    return "".join(out)


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(alphabet[i] for i in rng.integers(len(alphabet), size=n))


def ri(rng, lo, hi):
    """uniform integer in [lo, hi), tolerant of empty ranges"""
    return int(lo) if hi <= lo else int(rng.integers(lo, hi))


def sv_pair(rng, kind, n1, flank):
    """A reference window and a contig carrying one structural event."""
    w = rand_seq(rng, n1)
    if kind == "del":
        L = ri(rng, 5, max(6, n1 - 2 * flank))
        s = ri(rng, flank, n1 - flank - L + 1)
        c = w[s - flank:s] + w[s + L:s + L + flank]
    elif kind == "ins":
        p = ri(rng, flank, n1 - flank + 1)
        c = w[p - flank:p] + rand_seq(rng, ri(rng, 3, 40)) + w[p:p + flank]
    elif kind == "tdup":
        a = ri(rng, flank // 2, n1 // 2); b = ri(rng, a + 10, min(n1 - 5, a + 80))
        c = w[max(0, a - flank // 2):b] + w[a:min(n1, b + flank // 2)]
    elif kind == "inv":
        a = ri(rng, flank, n1 // 2); b = ri(rng, a + 20, n1 - flank)
        inv = w[:a] + revcom(w[a:b]) + w[b:]
        if rng.random() < 0.5:
            c = inv[max(0, a - flank):a + flank]
        else:
            c = inv[b - flank:min(n1, b + flank)]
    elif kind == "rep":   # low-complexity: many equivalent break points
        unit = rand_seq(rng, ri(rng, 1, 5))
        w = rand_seq(rng, flank) + unit * ri(rng, 6, 25) + rand_seq(rng, flank)
        c = w[:flank] + unit * ri(rng, 2, 12) + w[-flank:]
    else:  # unrelated
        c = rand_seq(rng, 2 * flank)
    return w, mutate(rng, c)


def decorate(rng, s, style):
    if style == "N":
        b = list(s)
        for _ in range(max(1, len(b) // 25)):
            b[int(rng.integers(len(b)))] = "NnRYxX"[int(rng.integers(6))]
        return "".join(b)
    if style == "lower":
        return "".join(ch.lower() if rng.random() < 0.5 else ch for ch in s)
    if style == "U":
        return "".join(("U" if rng.random() < 0.7 else "u") if (ch == "T" and rng.random() < 0.3) else ch for ch in s)
    return s


def random_cases(cases, n_cases=220, seed=20141):
    rng = np.random.default_rng(seed)
    scorings = [(1, -2, -2, -1), (1, -1, -10, -1), (2, -3, -5, -2), (1, -2, -4, -2), (3, -1, -1, -1), (1, -1, -1, -3)]
    modes = ["indel", "indel", "tdup", "inv", "invl", "invr"]
    kinds = {"indel": ["del", "ins", "rep", "none"], "tdup": ["tdup", "rep", "del"], "inv": ["inv", "del"],
             "invl": ["inv", "none"], "invr": ["inv", "rep"]}
    with tempfile.TemporaryDirectory() as td:
        for i in range(n_cases):
            mode = modes[i % len(modes)]
            kind = kinds[mode][int(rng.integers(len(kinds[mode])))]
            n1 = int(rng.integers(40, 420)); flank = int(rng.integers(8, max(9, min(60, n1 // 4))))
            w, c = sv_pair(rng, kind, n1, flank)
            style = ["plain", "plain", "N", "lower", "U"][int(rng.integers(5))]
            w, c = decorate(rng, w, style), decorate(rng, c, style)
            if rng.random() < 0.4:
                c = revcom(c)
            m, mm, go, ge = scorings[int(rng.integers(len(scorings)))]
            opts = [f"-{mode}", f"-match={m}", f"-mismatch={mm}", f"-go={go}", f"-ge={ge}"]
            if rng.random() < 0.5: opts.append("-both")
            if rng.random() < 0.15: opts.append("-revcom1")
            if rng.random() < 0.25: opts.append("-revcom2")
            if rng.random() < 0.2:
                a = int(rng.integers(1, max(2, len(w) // 4))); opts.append(f"-coor1={a}-{len(w) - int(rng.integers(0, 5))}")
            if rng.random() < 0.2:
                a = int(rng.integers(1, max(2, len(c) // 4))); opts.append(f"-coor2={a}-")
            Path(td, "a.fa").write_text(f">win{i} random window\n{w}\n")
            Path(td, "b.fa").write_text(f">ctg{i}\n{c}\n")
            cases.append(orig_case(f"rand{i}", "a.fa", "b.fa", opts, td))
    print(f"  {n_cases} randomised cases generated with the reference binary")


def cli_orig_fixture():
    """Original-dialect CLI fixture: multi-FASTA all-pairs + stdin command lines."""
    cli = OUT / "cli" / "orig"
    cli.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(77)
    w1, c1 = sv_pair(rng, "del", 300, 40)
    w2, c2 = sv_pair(rng, "inv", 260, 30)
    (cli / "first.fa").write_text(f">w1 first window\n{w1}\n>w2\n{w2[:120]}\n{w2[120:]}\n")
    (cli / "second.fa").write_text(f">c1\n{c1}\n>c2 contig two\n{c2}\n")
    cmds = [["-indel", "-both", "first.fa", "second.fa"],
            ["-inv", "-match=2", "-mismatch=-3", "-go=-5", "-ge=-2", "first.fa", "second.fa"],
            ["-tdup", "-revcom2", "-coor1=10-250", "first.fa", "second.fa"]]
    for k, c in enumerate(cmds):
        (cli / f"cmd{k}.txt").write_text(" ".join(c) + "\n")
        (cli / f"expected{k}.age").write_text(strip_time(run([BIN_ORIG] + c, cwd=cli)))
    print("  tests/golden/cli/orig written")


def main():
    if not BIN_BATCH.exists() or not BIN_ORIG.exists():
        sys.exit("run oracle/build_ref.sh first")
    OUT.mkdir(parents=True, exist_ok=True)
    cases = []
    reference_examples(cases)
    test1_example(cases)
    test2_example(cases)
    random_cases(cases)
    cli_orig_fixture()
    with open(OUT / "aligner_cases.json", "w") as f:
        json.dump({"generator": "oracle/make_golden.py", "cases": cases}, f, indent=0)
    print(f"wrote {len(cases)} aligner cases -> {OUT / 'aligner_cases.json'}")


if __name__ == "__main__":
    main()
