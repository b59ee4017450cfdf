#!/usr/bin/env python
"""Benchmark of the AGE realignment hot path (BASELINE.json: realignments/s and GCUPS vs host-core CPU AGE).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

A step is one pass of the hot path over one batch of B (default 7104) synthetic candidates of configuration C2
(5 kb reference windows, 2 x 500 bp contig flanks, -indel -both, scoring 1/-1/-10/-1).
Each rank (one process per GPU, torchrun) aligns its own shard -- candidates are independent, there is no
data-path collective -- so scaling is weak and `value` is the candidates all ranks finished / the max time.

  value        kernel-only realignments/s: inputs staged in HBM before the timed region (age_stage); the K steps are
               queued back to back (age_run_staged_many) and timed with CUDA events on the library's stream
  e2e          the same metric through the C ABI with HOST buffers (age_submit / age_collect, two steps in flight):
               packing, H2D, kernels, D2H, gather of every step inside the timed region; e2e.sync_value = one blocking
               age_align_batch call per step
  roofline     the dominant kernel (age_sweep_kernel) against the binding roofline, integer SIMD issue on the ALU pipe
               (peak = the age_int16_peak microbenchmark on this GPU); CUDA events around the kernel on the library's stream
  hbm          the secondary roofline of the same kernel: bytes it moves by construction / the same time / measured HBM peak
  extra        C4 (-inv, 20 kb x 4 kb) and C5 (-indel -both, 50 kb x 10 kb) kernel-only throughput
  inproc       a fixed C3-style mixed set (12 length bins, 70 % indel / 30 % tdup, two invocations) through ONE context
               created over all N devices (age_create(ids, N)): the product's own sharding; rank 0, after the per-rank legs
  cli          the age_align process, FASTA + variant table in -> .age text out (at N > 1: AGE_DEVICES=0..N-1)
  cpu_baseline the reference AGE binary on a bounded sample of the same workload, on the host cores; its stdout is
               diffed against the product CLI on the same files (parity_checked)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SCORING = (1, -1, -10, -1)          # AgeOption.h:25-28 (AsmvarDetect defaults) = Test2/Run.sh
SCORE_OPTS = ["-m", str(SCORING[0]), "-M", str(SCORING[1]), "-g", str(SCORING[2]), "-e", str(SCORING[3])]
# ALU-pipe instructions of the recurrence per cell update: 6 per pair cell in the forward sweep
# (VIMNMX, PRMT, VIADDMNMX.RELU, LOP3, VIADD, VIMNMX3), 8 in the reverse one (+ offset decode, VIADDMNMX.U16x2 for the split
# sum); a pair cell is two cell updates.  The two multiply-adds per pair cell run on the FMA pipe and are not counted.
ALU_INSTR_PER_CELL_UPDATE = (6 + 8) / 2 / 2
# DRAM traffic of age_sweep_kernel in the committed `ncu --set full` capture (profiles/r2u_sweep_key_metrics.txt):
# dram__bytes_read.sum 115.27 GB + dram__bytes_write.sum 121.26 GB for one launch over 164.63 G cell updates (7104 C2 candidates)
NCU_DRAM_BYTES_PER_CELL_UPDATE = (115.267648e9 + 121.262472e9) / 164.6323e9      # (round 1, 5920 candidates: the same 1.437 B per cell update)
WORKLOADS = {
    "c2": "C2 synthetic indel candidates: 5 kb windows, 2x500 bp flanks, -indel -both, 1/-1/-10/-1",
    "c4": "C4 synthetic inversions: 20 kb windows, 2 x 2 kb flanks, -inv, 1/-1/-10/-1",
    "c5": "C5 synthetic long deletions: 50 kb windows, 10 kb contigs, -indel -both, 1/-1/-10/-1",
}
INPROC_CANDIDATES = 98304            # C3-style set of the in-process multi-device leg, generated once ...
INPROC_REPEAT = 4                    # ... and aligned this many times over: 393 216 candidates at every N (strong scaling)


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------------
def both_jobs(cands, flag):
    from asmvar_b200 import synth
    import numpy as np
    jobs = []
    for w, q in cands:
        rc = synth.revcom(np.frombuffer(q, dtype=np.uint8)).tobytes()
        k = len(jobs)
        jobs.append((w, q, flag, *SCORING, k + 1))     # -both: the two jobs name each other (age_job.partner)
        jobs.append((w, rc, flag, *SCORING, k))
    return jobs


def workload_jobs(name: str, n: int, seed_shift: int = 0):
    """(candidates, jobs, aligners per candidate, CLI mode options)"""
    from asmvar_b200 import capi, synth
    if name == "c2":
        cands = synth.indel_candidates(n, seed=1002 + seed_shift)
        return cands, both_jobs(cands, capi.INDEL_FLAG), 2, ["-i", "-b"]
    if name == "c5":
        cands = synth.long_sv_candidates(n, seed=1005 + seed_shift)
        return cands, both_jobs(cands, capi.INDEL_FLAG), 2, ["-i", "-b"]
    cands = synth.inversion_candidates(n, seed=1004 + seed_shift)
    return cands, [(w, q, capi.INVERSION_FLAG, *SCORING, -1) for w, q in cands], 2, ["-v"]   # INVL + auxiliary INVR aligner


def cell_updates(cands, aligners: int = 2) -> float:
    return float(sum(2 * aligners * len(w) * len(q) for w, q in cands))   # forward + reverse, `aligners` per candidate


# ---------------------------------------------------------------------------------------------------------------
# the reference on the host cores (cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------------------------------------------
def run_cli_shards(exe: Path, cands, mode_opts, procs: int, workdir: Path, tag: str, env=None, keep_stdout: bool = True):
    """Run `exe` over `cands` split into `procs` shards, one process each, in parallel.  Returns (wall seconds,
    concatenated stdout in shard order, candidates in that order)."""
    from asmvar_b200 import synth
    if not exe.exists():
        raise FileNotFoundError(exe)
    ids = list(range(len(cands)))
    shards = [ids[i::procs] for i in range(procs) if ids[i::procs]]
    cmds, outs = [], []
    for i, sh in enumerate(shards):
        t, q, v = synth.write_files([cands[k] for k in sh], workdir, f"{tag}{i}", ids=sh)
        cmds.append([str(exe), *mode_opts, *SCORE_OPTS, "-t", str(t), "-q", str(q), str(v)])
        outs.append(workdir / f"{tag}{i}.age")
    t0 = time.perf_counter()
    running = []
    for c, o in zip(cmds, outs):
        fo = open(o, "wb") if keep_stdout else subprocess.DEVNULL
        running.append((subprocess.Popen(c, stdout=fo, stderr=subprocess.DEVNULL, env=env), fo))
    for p, fo in running:
        rc = p.wait()
        if keep_stdout:
            fo.close()
        if rc != 0:
            raise RuntimeError(f"{exe.name} failed ({rc})")
    dt = time.perf_counter() - t0
    text = b"".join(o.read_bytes() for o in outs) if keep_stdout else b""
    return dt, text, [k for sh in shards for k in sh]


def strip_time(text: bytes) -> bytes:
    return b"".join(l for l in text.splitlines(keepends=True) if not l.startswith(b"Alignment time is"))


def product_cli_text(cands, order, mode_opts, workdir: Path, tag: str, devices: str) -> bytes:
    """stdout of asmvar_b200/age_align over the same candidates (same record names), in `order`"""
    from asmvar_b200 import synth
    t, q, v = synth.write_files([cands[k] for k in order], workdir, tag, ids=order)
    env = dict(os.environ, AGE_DEVICES=devices)
    p = subprocess.run([str(ROOT / "asmvar_b200" / "age_align"), *mode_opts, *SCORE_OPTS, "-t", str(t), "-q", str(q), str(v)],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    if p.returncode != 0:
        raise RuntimeError("age_align failed: " + p.stderr.decode()[-500:])
    return p.stdout


def count_records(text: bytes) -> int:
    return sum(1 for l in text.splitlines() if l.startswith(b"# "))


def cpu_baseline_leg(cands, device: str, sample_n: int) -> dict:
    """The reference on the host cores over the first `sample_n` candidates of the step (-O2 build, one process per
    core), its as-shipped flavour (Makefile flags: no -O, OpenMP, two threads per alignment) on a quarter of them, and
    the diff of the reference's stdout against the product CLI on the same files."""
    cores = os.cpu_count() or 1
    ref = ROOT / "oracle" / "_ref"
    sample = cands[:sample_n]
    out = {"unit": "realignments/s", "cores": cores, "kind": "reference"}
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        d = Path(d)
        sec, text, order = run_cli_shards(ref / "age_align_ref", sample, ["-i", "-b"], cores, d, "r")
        out.update(value=len(sample) / sec, gcups=cell_updates(sample) / sec / 1e9, seconds=sec,
                   sample=f"first {len(sample)} candidates of the step, one age_align_ref (-O2, no OpenMP) process per core, {cores} cores")
        ours = product_cli_text(sample, order, ["-i", "-b"], d, "p", device)
        same = strip_time(ours) == strip_time(text)
        out["parity_checked"] = count_records(text) if same else 0
        out["parity"] = "identical" if same else "DIFFERENT"
        # as shipped: 2 OpenMP threads per process, so half as many processes as cores
        try:
            sub = sample[: max(cores, len(sample) // 4)]
            env = dict(os.environ, OMP_NUM_THREADS="2")
            sec2, _, _ = run_cli_shards(ref / "age_align_ref_shipped", sub, ["-i", "-b"], max(1, cores // 2), d, "s", env=env, keep_stdout=False)
            out["as_shipped"] = {"value": len(sub) / sec2, "gcups": cell_updates(sub) / sec2 / 1e9, "seconds": sec2,
                                 "sample": f"first {len(sub)} candidates, age_align_ref_shipped (the reference Makefile's flags: no -O, -fopenmp -DOMP, "
                                           f"2 threads per alignment), {max(1, cores // 2)} processes on {cores} cores"}
        except Exception as e:
            out["as_shipped"] = {"value": None, "sample": f"unavailable: {e}"}
    return out


def extra_parity(name: str, cands, mode_opts, device: str, n: int) -> dict:
    """reference stdout vs product CLI stdout on `n` candidates of an extra workload"""
    ref = ROOT / "oracle" / "_ref" / "age_align_ref"
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        d = Path(d)
        sec, text, order = run_cli_shards(ref, cands[:n], mode_opts, n, d, "r")
        ours = product_cli_text(cands[:n], order, mode_opts, d, "p", device)
        same = strip_time(ours) == strip_time(text)
    return {"parity_checked": count_records(text) if same else 0, "parity": "identical" if same else "DIFFERENT",
            "cpu_value": n / sec, "cpu_gcups": cell_updates(cands[:n]) / sec / 1e9, "cpu_processes": n}


def reference_arm(args, rank: int):
    """--impl reference: the reference's own CPU implementation on the box's host cores (rank 0 only)."""
    if rank != 0:
        return
    from asmvar_b200 import synth
    cores = os.cpu_count() or 1
    per_step = max(cores, 8) * 4          # ~10 s of CPU work per step on 16 cores
    cands = synth.indel_candidates(per_step, seed=1002)
    cu = cell_updates(cands)
    exe = ROOT / "oracle" / "_ref" / "age_align_ref"
    with tempfile.TemporaryDirectory() as d:
        for _ in range(args.warmup):
            run_cli_shards(exe, cands[:cores], ["-i", "-b"], cores, Path(d), "w", keep_stdout=False)
        total = sum(run_cli_shards(exe, cands, ["-i", "-b"], cores, Path(d), "r", keep_stdout=False)[0] for _ in range(args.steps))
    v = per_step * args.steps / total
    print(json.dumps({
        "impl": "reference", "metric": "AGE realignments/sec", "value": v, "unit": "realignments/s", "gcups": cu * args.steps / total / 1e9,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": {"workload": WORKLOADS["c2"]},
        "sizes": {"candidates_per_step": per_step},
        "cpu_baseline": {"value": v, "unit": "realignments/s", "cores": cores, "kind": "reference",
                         "sample": f"{per_step} C2 candidates per step, one age_align_ref (-O2, no OpenMP) process per core"},
        "e2e": {"value": v, "unit": "realignments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# GPU legs
# ---------------------------------------------------------------------------------------------------------------
def kernel_leg(eng, arr, n, steps, warmup, barrier, sampler=None):
    """K steps of the staged batch queued back to back; returns (total ms by CUDA events, launches)"""
    eng.stage(arr, n)
    eng.run_staged_many(max(1, warmup))
    barrier()
    if sampler:
        sampler.start()
    l0 = eng.stats()["launches"]
    total = eng.run_staged_many(steps)
    return total, int(eng.stats()["launches"] - l0)


def sweep_kernel_leg(eng, arr, n, steps):
    """The dominant kernel by itself: one age_run_staged per step, CUDA events around the sweep kernel(s) on the library's
    stream (age_stats.fill_ms).  Returns (mean sweep-kernel ms, mean ms of all kernels of a step, sweep bytes of the batch)."""
    eng.stage(arr, n)
    eng.run_staged()
    fill = k = 0.0
    for _ in range(steps):
        eng.run_staged()
        st = eng.stats()
        fill += st["fill_ms"]; k += st["kernel_ms"]
    eng.fetch(n)
    return fill / steps, k / steps, eng.stats()["sweep_bytes"]


def inproc_leg(devices, single_device_results=None):
    """A fixed C3-style set through one context over `devices` (the product's own multi-device path): 12 length bins,
    70 % indel / 30 % tdup, run as two invocations (one mode per invocation, age_align.cpp:207-221), streamed in batches."""
    from asmvar_b200 import capi, engine, synth
    t0 = time.perf_counter()
    mixed = synth.mixed_candidates(INPROC_CANDIDATES, seed=1003)
    gen_s = time.perf_counter() - t0
    kinds = {"indel": capi.INDEL_FLAG, "tdup": capi.TDUPLICATION_FLAG}
    nd = len(devices)
    per_batch = 4096 * nd
    summary = {}
    total_s = 0.0
    total_c = 0
    total_cu = 0.0
    checks = []
    with engine.Engine(list(devices)) as eng:
        for kind, flag in kinds.items():
            cands = [(w, q) for k, w, q in mixed if k == kind]
            # length-binned: candidates of similar DP area share a batch, so no warp is left with one long pair while the
            # others run out of short ones (the caller owns the order of its results; nothing is printed here)
            cands.sort(key=lambda c: len(c[0]) * len(c[1]), reverse=True)
            batches = [cands[i:i + per_batch] for i in range(0, len(cands), per_batch)]
            prepared = [eng.make_jobs(both_jobs(b, flag)) for b in batches] * INPROC_REPEAT      # the same job arrays again: nothing is cached between batches
            n_max = max(p[2] for p in prepared)
            outs = [(capi.AgeResult * max(1, n_max))() for _ in range(2)]          # batches differ in size: two result arrays of the largest
            # warm-up: first batch once (arena allocation, clocks)
            eng.align_raw(prepared[0][0], prepared[0][2], outs[0])
            t0 = time.perf_counter()
            pending = []
            scores = []
            reruns = 0
            for bi, (arr, keep, n) in enumerate(prepared):
                eng.submit(arr, n)
                pending.append((bi, n))
                if len(pending) == 2:
                    b0, n0 = pending.pop(0)
                    res = eng.collect(n0, outs[b0 & 1])
                    reruns += eng.stats()["pool_reruns"]
                    if b0 == 0:
                        scores = [(res[i].status, res[i].score, res[i].n_bp, res[i].detail) for i in range(min(n0, 2048))]
            while pending:
                b0, n0 = pending.pop(0)
                res = eng.collect(n0, outs[b0 & 1])
                reruns += eng.stats()["pool_reruns"]
                if b0 == 0:
                    scores = [(res[i].status, res[i].score, res[i].n_bp, res[i].detail) for i in range(min(n0, 2048))]
            dt = time.perf_counter() - t0
            total_s += dt; total_c += len(cands) * INPROC_REPEAT; total_cu += cell_updates(cands) * INPROC_REPEAT
            summary[kind] = {"candidates": len(cands) * INPROC_REPEAT, "seconds": dt, "value": len(cands) * INPROC_REPEAT / dt, "batches": len(prepared), "pool_reruns": int(reruns)}
            checks.append((kind, scores))
    return {"value": total_c / total_s, "unit": "realignments/s", "gcups": total_cu / total_s / 1e9, "candidates": total_c, "seconds": total_s,
            "devices": nd, "per_mode": summary, "generate_s": gen_s, "scaling": "strong",
            "api": f"one age_create over {nd} device(s); age_submit/age_collect, length-binned batches of {per_batch} candidates, two in flight",
            "workload": f"C3-style fixed set ({INPROC_CANDIDATES} candidates x {INPROC_REPEAT}): len1 in {{1,2,4,8}} kb x len2 in {{0.5,1,2}} kb, 70 % indel / 30 % tdup, -both"}, checks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0, help="candidates per step per GPU (default 7104 for C2 = 3 pairs per resident warp: 148 SMs x 16 warps, about 125 GB of scratch; "
                                                          "296 for C4, 74 for C5 = one cluster of two SMs per pair)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-cli", action="store_true", help="skip the age_align process leg (FASTA in -> .age out)")
    ap.add_argument("--no-extra", action="store_true", help="skip the C4 / C5 legs")
    ap.add_argument("--no-inproc", action="store_true", help="skip the in-process multi-device leg")
    ap.add_argument("--cli-candidates", type=int, default=100640, help="candidates of the age_align process leg (C2 asks for N = 100 000)")
    ap.add_argument("--cpu-sample", type=int, default=1024, help="candidates of the cpu_baseline leg")
    ap.add_argument("--workload", default="c2", choices=["c2", "c4", "c5"],
                    help="c2 (default, the configuration the metric is quoted on); c4 = -inv 20 kb x 4 kb; c5 = 50 kb x 10 kb -indel -both")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the AGE realigner has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as g
    if rank == 0:
        g.build(quiet=True)
    if world > 1:
        dist.barrier()
    from asmvar_b200 import capi, engine

    if args.batch <= 0:
        args.batch = {"c2": 7104, "c4": 296, "c5": 74}[args.workload]
    cands, jobs, aligners, mode_opts = workload_jobs(args.workload, args.batch, rank)
    cu = cell_updates(cands, aligners)
    eng = engine.Engine([local])
    arr, keep, n = eng.make_jobs(jobs)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- kernel-only: inputs resident in HBM, K steps queued back to back --------------------------------------
    sampler = ClockSampler(local)
    k_ms, launches = kernel_leg(eng, arr, n, args.steps, args.warmup, barrier, sampler if rank == 0 else None)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    res = eng.fetch(n)
    bad = sum(1 for i in range(n) if res[i].status != 0)
    if bad:
        raise SystemExit(f"bench.py: {bad} jobs failed")

    # ---- end to end: host buffers in, results out ---------------------------------------------------------
    # (a) streaming form of the C ABI (age_submit / age_collect, what the age_align CLI uses): step k+1 is packed and
    #     copied to the device while the kernels of step k run; every step's inputs come from host memory and every
    #     step's results are read back and assembled inside the timed region.
    out = (capi.AgeResult * n)()              # the caller's result buffer (host memory), reused across steps
    def stream(steps):
        eng.submit(arr, n)
        for _ in range(steps - 1):
            eng.submit(arr, n)
            eng.collect(n, out)
        eng.collect(n, out)
    stream(max(2, args.warmup - 1))
    barrier()
    t0 = time.perf_counter()
    stream(args.steps)
    st = eng.stats()
    barrier()
    e2e_s = time.perf_counter() - t0
    # (b) one synchronous age_align_batch call per step (nothing overlaps between steps)
    eng.align_raw(arr, n, out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.align_raw(arr, n, out)
    barrier()
    sync_s = time.perf_counter() - t0

    # ---- the dominant kernel alone (roofline) ---------------------------------------------------------------------
    fill_ms, serial_ms, sweep_bytes = sweep_kernel_leg(eng, arr, n, min(args.steps, 4))

    t = torch.tensor([k_ms, e2e_s * 1e3, sync_s * 1e3, fill_ms, serial_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    k_ms, e2e_ms, sync_ms, fill_ms, serial_ms = [float(x) for x in t.tolist()]

    line = None
    if rank == 0:
        total_c = args.batch * world * args.steps
        total_cu = cu * world * args.steps            # every rank draws the same distribution
        hbm, which = measured_peaks()
        line = {
            "metric": "AGE realignments/sec", "value": total_c / (k_ms / 1e3), "unit": "realignments/s", "gcups": total_cu / (k_ms / 1e3) / 1e9,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": k_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]},
            "sizes": {"candidates_per_step_per_gpu": args.batch, "aligners_per_candidate": aligners,
                      "l2": "per-step working set (sequences + Fmax stream, > 100 GB) exceeds the 126 MB L2"},
            "clocks": clocks,
            "e2e": {"value": total_c / (e2e_ms / 1e3), "unit": "realignments/s", "gcups": total_cu / (e2e_ms / 1e3) / 1e9,
                    "h2d_bytes_per_step": st["h2d_bytes"], "d2h_bytes_per_step": st["d2h_bytes"],
                    "api": "age_submit/age_collect, two steps in flight",
                    "sync_value": total_c / (sync_ms / 1e3), "sync_api": "one age_align_batch call per step, no overlap between steps"},
            "gpu_launches": launches,
            "sync_ms_per_step": serial_ms,
        }
        peak = C.c_double(0.0)
        ach_i = ALU_INSTR_PER_CELL_UPDATE * cu / (fill_ms / 1e3) / 1e9
        have_peak = capi.load().age_int16_peak(local, C.byref(peak)) == 0 and peak.value > 0
        traffic = NCU_DRAM_BYTES_PER_CELL_UPDATE * cu / 1e9 if args.workload == "c2" else None
        line["roofline"] = {
            "bound": "int16x2 SIMD issue (ALU pipe)", "achieved": ach_i, "peak": peak.value if have_peak else None, "unit": "G lane-instr/s",
            "frac": ach_i / peak.value if have_peak else None, "traffic": traffic, "traffic_unit": "GB of DRAM read+write per launch (dram__bytes of the ncu capture in profiles/r2u_sweep_key_metrics.txt per cell update, times the cell updates of this launch)",
            "kernel": "age_sweep_kernel", "kernel_ms": fill_ms, "cell_updates_per_launch": cu, "alu_instr_per_cell_update": ALU_INSTR_PER_CELL_UPDATE,
            "note": "achieved = 3.5 ALU-pipe instructions per cell update (6 forward / 8 reverse per packed pair cell; the FMA-pipe multiply-adds are not "
                    "counted) x cell updates of one launch / the kernel's CUDA-event time; peak = age_int16_peak microbenchmark on this GPU "
                    "(independent chains of the same ALU instruction mix on every SM)"}
        line["hbm"] = {"bound": "hbm", "achieved": sweep_bytes / (fill_ms / 1e3) / 1e9, "peak": hbm, "unit": "GB/s", "peak_source": which,
                       "frac": sweep_bytes / (fill_ms / 1e3) / 1e9 / hbm, "bytes_per_cell_update": sweep_bytes / cu,
                       "ncu_dram_bytes_per_cell_update": NCU_DRAM_BYTES_PER_CELL_UPDATE if args.workload == "c2" else None,
                       "note": "secondary roofline of the same kernel: bytes it moves by construction for this batch (D planes in the batch's storage "
                               "format written and read back, checkpoints, boundary rows, tile maxima; age_stats.sweep_bytes) / the same time"}
    eng.close()                               # the legs below need the HBM (other processes / a multi-device context)
    eng = None

    if rank == 0 and not args.no_extra and args.workload == "c2":
        extra = {}
        for name, batch in (("c4", 296), ("c5", 74)):
            try:
                xc, xj, xa, xm = workload_jobs(name, batch)
                with engine.Engine([local]) as xe:
                    xarr, xkeep, xn = xe.make_jobs(xj)
                    ms, _ = kernel_leg(xe, xarr, xn, 4, 2, lambda: torch.cuda.synchronize())
                    xs = xe.stats()
                    xres = xe.fetch(xn)
                    ok = all(xres[i].status == 0 for i in range(xn))
                xcu = cell_updates(xc, xa)
                extra[name] = {"workload": WORKLOADS[name], "candidates_per_step": batch, "value": batch * 4 / (ms / 1e3), "unit": "realignments/s",
                               "gcups": xcu * 4 / (ms / 1e3) / 1e9, "ms_per_step": ms / 4, "all_ok": ok,
                               "int_frac": (ALU_INSTR_PER_CELL_UPDATE * xcu * 4 / (ms / 1e3) / 1e9 / peak.value) if have_peak else None,
                               "int_frac_note": "whole step (sweeps + pass 2) against the ALU-pipe peak"}
                if not args.no_cpu:
                    extra[name].update(extra_parity(name, xc, xm, str(local), 8 if name == "c4" else 4))      # reference stdout diffed on 8 C4 / 4 C5 candidates
            except Exception as e:
                extra[name] = {"value": None, "note": f"unavailable: {e}"}
        line["extra"] = extra

    cpu_group = None
    if world > 1:
        dist.barrier()                        # every rank has released its context
        cpu_group = dist.new_group(backend="gloo")   # the other ranks wait for rank 0's legs on the CPU, not in a spinning NCCL kernel
    if rank == 0 and not args.no_inproc and args.workload == "c2":
        try:
            line["inproc"], checks = inproc_leg(list(range(world)))
            if world > 1:                     # the same first candidates on one device: results must be identical
                line["inproc"]["identical_to_single_device"] = checks == inproc_check_single(local)
        except Exception as e:
            line["inproc"] = {"value": None, "note": f"unavailable: {e}"}
    if rank == 0 and not args.no_cli and args.workload == "c2":
        try:
            sys.path.insert(0, str(ROOT / "tools"))
            import cli_throughput
            devs = ",".join(str(i) for i in range(world))
            m = cli_throughput.measure(args.cli_candidates, devices=devs, replicate=world, also_binary=True)      # N x the candidates at N devices
            mb = m["binary"]
            line["cli_binary"] = {"value": mb["realign_per_s"], "unit": "realignments/s", "candidates": mb["candidates"], "seconds": mb["seconds"],
                                  "record_bytes": mb["age_bytes"], "steady_value": mb.get("steady_realign_per_s"), "devices": devs,
                                  "note": "the same process with --binary-out: compact records instead of text (age2text prints the text from them later)"}
            line["cli"] = {"value": m["realign_per_s"], "unit": "realignments/s", "candidates": m["candidates"], "seconds": m["seconds"],
                           "fasta_bytes": m["fasta_bytes"], "age_bytes": m["age_bytes"], "cuda_startup_s": m.get("cuda_startup_s"),
                           "steady_value": m.get("steady_realign_per_s"), "devices": devs,
                           "note": "whole age_align process, FASTA + variant table in -> .age text out (SURVEY 8d end-to-end), files in /dev/shm; "
                                   "includes process and CUDA start-up; steady_value = candidates / time after the device context is ready"}
        except Exception as e:
            line["cli"] = {"value": None, "note": f"unavailable: {e}"}
    if rank == 0 and world == 1 and not args.no_cpu and args.workload == "c2":
        try:
            line["cpu_baseline"] = cpu_baseline_leg(cands, str(local), min(len(cands), args.cpu_sample))
        except Exception as e:   # the baseline is reported, never required for the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "realignments/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(group=cpu_group)
        dist.destroy_process_group()


def inproc_check_single(device: int):
    """the first 1024 candidates of each mode of the in-process leg on ONE device, for the identity check at N > 1"""
    from asmvar_b200 import capi, engine, synth
    mixed = synth.mixed_candidates(INPROC_CANDIDATES, seed=1003)
    checks = []
    with engine.Engine([device]) as eng:
        for kind, flag in (("indel", capi.INDEL_FLAG), ("tdup", capi.TDUPLICATION_FLAG)):
            cands = [(w, q) for k, w, q in mixed if k == kind]
            cands.sort(key=lambda c: len(c[0]) * len(c[1]), reverse=True)
            arr, keep, n = eng.make_jobs(both_jobs(cands[:1024], flag))
            res = eng.align_raw(arr, n)
            checks.append((kind, [(res[i].status, res[i].score, res[i].n_bp, res[i].detail) for i in range(min(n, 2048))]))
    return checks


if __name__ == "__main__":
    main()
