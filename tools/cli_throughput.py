#!/usr/bin/env python
"""FASTA in -> .age out throughput of asmvar_b200/age_align (SURVEY 8(d) end-to-end figure) on C2 candidates."""
import argparse, os, subprocess, sys, tempfile, time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def run_once(cmd, out, env, n, to_devnull, verbose):
    t0 = time.perf_counter()
    with open(os.devnull if to_devnull else out, "wb") as fo:
        p = subprocess.run(cmd, stdout=fo, stderr=subprocess.PIPE, env=env)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError(p.stderr.decode()[-2000:])
    ready = None
    for line in p.stderr.decode().splitlines():
        if line.startswith("[cli prof] device context ready"):
            ready = float(line.split(" at ")[1].split()[0])
        if verbose and line.startswith(("[cli prof]", "[host prof]")):
            print(line, file=sys.stderr)
    res = {"candidates": n, "seconds": dt, "realign_per_s": n / dt, "age_bytes": os.path.getsize(out)}
    if ready is not None and dt > ready:
        res["cuda_startup_s"] = ready
        res["steady_realign_per_s"] = n / (dt - ready)
    return res


def measure(n: int, devices: str = "0", chunk: int = 0, keep: str = "", verbose: bool = False, replicate: int = 1, also_binary: bool = False) -> dict:
    """`n` C2 candidates, written `replicate` times under different record names (several GPUs need more input than is
    worth generating in Python).  With also_binary the same files go through --binary-out as well (result["binary"])."""
    from asmvar_b200 import synth
    cands = synth.indel_candidates(n, seed=1002)
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        t, q, v = synth.write_files(cands, Path(d), "cli")
        if replicate > 1:
            with open(t, "ab") as ft, open(q, "ab") as fq, open(v, "a") as fv:
                for rep in range(1, replicate):
                    for i, (w, c) in enumerate(cands):
                        ft.write(b">t%d_%d\n%s\n" % (i, rep, w))
                        fq.write(b">q%d_%d\n%s\n" % (i, rep, c))
                        fv.write(f"t{i}_{rep}\t1\t{len(w)}\tq{i}_{rep}\t1\t{len(c)}\tcand{i}_{rep}\n")
            n = n * replicate
        env = dict(os.environ, AGE_DEVICES=devices, AGE_CLI_PROF="1")
        if chunk:
            env["AGE_CHUNK"] = str(chunk)
        out = keep or os.path.join(d, "out.age")
        cmd = [str(ROOT / "asmvar_b200" / "age_align"), "-i", "-b", "-m", "1", "-M", "-1", "-g", "-10", "-e", "-1", "-t", str(t), "-q", str(q), str(v)]
        res = run_once(cmd, out, env, n, False, verbose)
        res["fasta_bytes"] = sum(os.path.getsize(x) for x in (t, q, v))
        if also_binary:
            time.sleep(2.0)      # the driver is still tearing down the first process's > 100 GB context: a start right behind it waits for that
            agb = os.path.join(d, "out.agb")
            res["binary"] = run_once(cmd[:1] + ["--binary-out", agb] + cmd[1:], agb, env, n, True, verbose)
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-n", type=int, default=14208)
    ap.add_argument("--devices", default="0")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--binary", action="store_true")
    a = ap.parse_args()
    print(measure(a.n, a.devices, a.chunk, verbose=True, also_binary=a.binary))
