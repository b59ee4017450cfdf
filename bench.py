#!/usr/bin/env python
"""Benchmark of the AGE realignment hot path (BASELINE.json: realignments/s and GCUPS vs host-core CPU AGE).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

A step is one pass of the hot path over one batch of B synthetic candidates of configuration C2
(5 kb reference windows, 2 x 500 bp contig flanks, -indel -both, scoring 1/-1/-10/-1).
Each rank (one process per GPU, torchrun) aligns its own shard -- candidates are independent, there is no
data-path collective -- so scaling is weak and `value` is the candidates all ranks finished / the max time.

  value        kernel-only realignments/s: inputs staged in HBM before the timed region (age_stage), each step
               = age_run_staged (all kernels of the batch), timed with CUDA events on the library's stream
  e2e          the same metric through the C ABI with HOST buffers (age_submit / age_collect, two steps in flight):
               packing, H2D, kernels, D2H, gather of every step inside the timed region; e2e.sync_value = one blocking
               age_align_batch call per step
  cli          the age_align process, FASTA + variant table in -> .age text out
  roofline     the dominant kernel (age_sweep_kernel) against the measured HBM copy peak, algorithmic bytes
  int_roofline the same kernel against the measured int16x2 SIMD issue peak (age_int16_peak microbenchmark)
  cpu_baseline the reference AGE binary (oracle/_ref/age_align_ref, -O2, one process per host core) on a bounded
               sample of the same workload
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
ALG_BYTES_PER_CELL_UPDATE = 2.5     # indel = 5 B per matrix cell = 2.5 B per cell update
# DRAM traffic of age_sweep_kernel from the committed `ncu --set full` capture (profiles/r1b_sweep_kernel_ncu_key_metrics.txt):
# dram__bytes_read.sum 96.09 GB + dram__bytes_write.sum 101.08 GB for one launch over 137.24 G cell updates (5920 candidates)
NCU_DRAM_BYTES_PER_CELL_UPDATE = (96.086718e9 + 101.077318e9) / 137.23574e9
KERNELS_PER_STEP = 6                # per age_run_staged: age_sweep, age_select, age_presweep x2, age_enum, age_blocks (age_pack runs at stage time;
                                    # an end-to-end step launches 7)
# s16x2 / logic instructions of the recurrence per cell pair: VIMNMX, PRMT, VIADDMNMX, LOP3, VIADD,
# tag clearing, VIMNMX3 in both sweeps, + offset decode and VIADDMNMX.U16x2 for the split sum in the reverse sweep
# => (7 + 9) / 2 per pair-cell visit, and a pair-cell visit is two cell updates
SIMD_INSTR_PER_CELL_UPDATE = (7 + 9) / 2 / 2
WORKLOAD = "C2 synthetic indel candidates: 5 kb windows, 2x500 bp flanks, -indel -both, 1/-1/-10/-1"


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
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
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


def make_jobs(cands, flag):
    from asmvar_b200 import synth
    import numpy as np
    jobs = []
    for w, q in cands:
        rc = synth.revcom(np.frombuffer(q, dtype=np.uint8)).tobytes()
        k = len(jobs)
        jobs.append((w, q, flag, *SCORING, k + 1))     # -both: the two jobs name each other (age_job.partner)
        jobs.append((w, rc, flag, *SCORING, k))
    return jobs


def cell_updates(cands) -> float:
    return float(sum(2 * 2 * len(w) * len(q) for w, q in cands))   # fwd+rev, two aligners per -both candidate


def run_reference(cands, cores: int, workdir: Path) -> float:
    """Wall seconds of oracle/_ref/age_align_ref over `cands`, one process per core (non-OpenMP -O2 build)."""
    from asmvar_b200 import synth
    exe = ROOT / "oracle" / "_ref" / "age_align_ref"
    if not exe.exists():
        raise FileNotFoundError(exe)
    shards = [cands[i::cores] for i in range(cores) if cands[i::cores]]
    cmds = []
    for i, sh in enumerate(shards):
        t, q, v = synth.write_files(sh, workdir, f"s{i}")
        cmds.append([str(exe), "-i", "-b", "-m", str(SCORING[0]), "-M", str(SCORING[1]), "-g", str(SCORING[2]), "-e", str(SCORING[3]),
                     "-t", str(t), "-q", str(q), str(v)])
    t0 = time.perf_counter()
    procs = [subprocess.Popen(c, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for c in cmds]
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("reference age_align failed")
    return time.perf_counter() - t0


def reference_arm(args, rank: int):
    """--impl reference: the reference's own CPU implementation on the box's host cores (rank 0 only)."""
    if rank != 0:
        return
    from asmvar_b200 import synth
    cores = os.cpu_count() or 1
    per_step = max(cores, 8) * 4          # ~10 s of CPU work per step on the box's 16 cores
    total = None
    cands = synth.indel_candidates(per_step, seed=1002)
    cu = cell_updates(cands)
    with tempfile.TemporaryDirectory() as d:
        for _ in range(args.warmup):
            run_reference(cands[:cores], cores, Path(d))
        total = sum(run_reference(cands, cores, Path(d)) for _ in range(args.steps))
    v = per_step * args.steps / total
    print(json.dumps({
        "impl": "reference", "metric": "AGE realignments/sec", "value": v, "unit": "realignments/s", "gcups": cu * args.steps / total / 1e9,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "candidates_per_step": per_step},
        "cpu_baseline": {"value": v, "unit": "realignments/s", "cores": cores, "kind": "reference",
                         "sample": f"{per_step} C2 candidates per step, one age_align_ref (-O2, no OpenMP) process per core"},
        "e2e": {"value": v, "unit": "realignments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=5920, help="candidates per step per GPU (5920 = 2.5 pairs per resident warp: 148 SMs x 16 warps; about 137 GB of scratch; the longest-first deal needs more pairs than warps to balance)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-cli", action="store_true", help="skip the age_align process leg (FASTA in -> .age out)")
    ap.add_argument("--cli-candidates", type=int, default=100640, help="candidates of the age_align process leg (C2 asks for N = 100 000; 17 chunks of 5920)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c4", "c5"],
                    help="c2 (default, the configuration the metric is quoted on); c4 = -inv 20 kb x 4 kb; c5 = 50 kb x 10 kb -indel -both "
                         "(extra measurements, not the headline)")
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
    from asmvar_b200 import capi, engine, synth

    global WORKLOAD
    if args.workload == "c2":
        cands = synth.indel_candidates(args.batch, seed=1002 + rank)
        jobs = make_jobs(cands, capi.INDEL_FLAG)
    elif args.workload == "c5":
        WORKLOAD = "C5 synthetic long deletions: 50 kb windows, 10 kb contigs, -indel -both, 1/-1/-10/-1"
        cands = synth.long_sv_candidates(args.batch, seed=1005 + rank)
        jobs = make_jobs(cands, capi.INDEL_FLAG)
    else:
        WORKLOAD = "C4 synthetic inversions: 20 kb windows, 2 x 2 kb flanks, -inv, 1/-1/-10/-1"
        cands = synth.inversion_candidates(args.batch, seed=1004 + rank)
        jobs = [(w, q, capi.INVERSION_FLAG, *SCORING, -1) for w, q in cands]      # INVL + auxiliary INVR aligner per candidate
    cu = cell_updates(cands)
    eng = engine.Engine([local])
    arr, keep, n = eng.make_jobs(jobs)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- kernel-only: inputs resident in HBM ------------------------------------------------------------
    eng.stage(arr, n)
    for _ in range(args.warmup):
        eng.run_staged()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    k_ms = fill_ms = 0.0
    launches = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.run_staged()
        st = eng.stats()
        k_ms += st["kernel_ms"]; fill_ms += st["fill_ms"]
    barrier()
    wall_kernel = time.perf_counter() - t0
    launches = args.steps * KERNELS_PER_STEP
    clocks = sampler.stop() if rank == 0 else None
    res = eng.fetch(n)
    bad = sum(1 for i in range(n) if res[i].status != 0)
    if bad and not os.environ.get("AGE_EXPERIMENT_FAST"):
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
    # (b) one synchronous age_align_batch call per step (nothing overlaps)
    eng.align_raw(arr, n, out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.align_raw(arr, n, out)
    barrier()
    sync_s = time.perf_counter() - t0

    t = torch.tensor([k_ms, fill_ms, e2e_s * 1e3, wall_kernel * 1e3, sync_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    k_ms, fill_ms, e2e_ms, wall_ms, sync_ms = [float(x) for x in t.tolist()]

    if rank == 0:
        total_c = args.batch * world * args.steps
        total_cu = cu * world * args.steps            # every rank draws the same distribution
        value = total_c / (k_ms / 1e3)
        hbm, which = measured_peaks()
        ach = ALG_BYTES_PER_CELL_UPDATE * cu * args.steps / (fill_ms / 1e3) / 1e9
        traffic = NCU_DRAM_BYTES_PER_CELL_UPDATE * cu / 1e9 if args.workload == "c2" else None
        dram_ach = None if traffic is None else traffic * args.steps / (fill_ms / 1e3)
        out = {
            "metric": "AGE realignments/sec", "value": value, "unit": "realignments/s", "gcups": total_cu / (k_ms / 1e3) / 1e9,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": k_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "candidates_per_step_per_gpu": args.batch, "aligners_per_candidate": 2,
                       "l2": "per-step working set (sequences + Fmax stream) exceeds the 126 MB L2"},
            "clocks": clocks,
            "e2e": {"value": total_c / (e2e_ms / 1e3), "unit": "realignments/s", "gcups": total_cu / (e2e_ms / 1e3) / 1e9,
                    "h2d_bytes_per_step": st["h2d_bytes"], "d2h_bytes_per_step": st["d2h_bytes"],
                    "api": "age_submit/age_collect, two steps in flight",
                    "sync_value": total_c / (sync_ms / 1e3), "sync_api": "one age_align_batch call per step, no overlap"},
            "gpu_launches": launches,
            "kernel_wall_ms_per_step": wall_ms / args.steps,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                         "traffic": traffic, "traffic_unit": "GB per launch",
                         "dram_achieved": dram_ach, "dram_frac": None if dram_ach is None else dram_ach / hbm,
                         "kernel": "age_sweep_kernel", "peak_source": which,
                         "note": "achieved = 2.5 algorithmic B per cell update (SURVEY 8d: 16-bit Fmax written + read, 1 trace byte per cell) x cell "
                                 "updates of one launch / its CUDA-event time, per GPU.  frac can exceed 1: this design moves fewer bytes than that "
                                 "model (no trace in the main sweep, Fmax as 8-bit row offsets): traffic = ncu dram read+write of the committed "
                                 "capture (1.44 B per cell update) scaled to this launch, dram_achieved = traffic / the same time.  The kernel is "
                                 "bound by integer SIMD issue (int_roofline, ALU pipe 72 % active in the capture)"},
        }
        peak = C.c_double(0.0)
        if capi.load().age_int16_peak(local, C.byref(peak)) == 0 and peak.value > 0:
            ach_i = SIMD_INSTR_PER_CELL_UPDATE * cu * args.steps / (fill_ms / 1e3) / 1e9
            out["int_roofline"] = {"bound": "int16x2 SIMD issue", "achieved": ach_i, "peak": peak.value, "unit": "G lane-instr/s",
                                   "frac": ach_i / peak.value, "kernel": "age_sweep_kernel",
                                   "note": "achieved = 4 algorithmic s16x2/logic instructions per cell update x cell updates / sweep-kernel time; "
                                           "peak = age_int16_peak microbenchmark on this GPU (independent chains of the same instruction mix)"}
        eng.close()                               # the CLI leg below is a separate process and needs the HBM
        eng = None
        if world == 1 and not args.no_cli and args.workload == "c2":
            try:
                sys.path.insert(0, str(ROOT / "tools"))
                import cli_throughput
                m = cli_throughput.measure(args.cli_candidates, devices=str(local))
                out["cli"] = {"value": m["realign_per_s"], "unit": "realignments/s", "candidates": m["candidates"], "seconds": m["seconds"],
                              "fasta_bytes": m["fasta_bytes"], "age_bytes": m["age_bytes"], "cuda_startup_s": m.get("cuda_startup_s"),
                              "steady_value": m.get("steady_realign_per_s"),
                              "note": "whole age_align process, FASTA + variant table in -> .age text out (SURVEY 8d end-to-end), files in /dev/shm; "
                                      "includes process and CUDA start-up; steady_value = candidates / time after the device context is ready"}
            except Exception as e:
                out["cli"] = {"value": None, "note": f"unavailable: {e}"}
        if world == 1 and not args.no_cpu and args.workload == "c2":
            try:
                cores = os.cpu_count() or 1
                sample = cands[: min(len(cands), max(cores, 8) * 8)]      # ~20 s of CPU work, a bit over a second of wall time
                with tempfile.TemporaryDirectory() as d:
                    sec = run_reference(sample, cores, Path(d))
                out["cpu_baseline"] = {"value": len(sample) / sec, "unit": "realignments/s", "gcups": cell_updates(sample) / sec / 1e9,
                                       "cores": cores, "kind": "reference",
                                       "sample": f"first {len(sample)} candidates of the step, one age_align_ref (-O2, no OpenMP) process per core"}
            except Exception as e:   # the baseline is reported, never required for the GPU number
                out["cpu_baseline"] = {"value": None, "unit": "realignments/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
        print(json.dumps(out), flush=True)
    if eng is not None:
        eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
