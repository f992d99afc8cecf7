#!/usr/bin/env python
"""bench.py — ORB build benchmark (BASELINE.json metric: build time & particle-passes/s vs HBM roofline).

A "step" is one full ORB build (count-left + bisection per level, partition, next axis) of the
workload through the C ABI (liborb_b200.so).  N=1 workload = BASELINE config[1]: 2^24 uniform
particles (reference generator), 2^12 leaf cells, reference-compatible level count.  N>1: every
rank holds its own 2^24-particle slice of the single generator stream (weak scaling), only the
per-cell count vectors cross NVLink (NCCL allreduce inside the library).

value      = reference-equivalent particle-passes per second, whole job, particles resident in HBM:
             sum over bisection iterations of particles in still-unfound cells (what orbit.cpp's
             loop streams; for the same particles identical for the reference and for this build,
             because the cut sequence is bit-identical) / build time.  (The reference arm counts its
             own passes: run on several threads the unmodified reference draws a different, partly
             duplicated particle set - its generator state is shared by its threads, init.cu:11-25 -
             so its numerator differs by some percent; each arm divides its own work by its own time.)
ms_per_step= ORB build time (CUDA events on the library's stream, max over ranks).
e2e        = same metric through the public API with HOST buffers: pinned host x,y,z -> device,
             build, device -> host x,y,z + cell heap + ranges, inside the timed region.
roofline   = dominant kernel (largest share of the step), algorithmic bytes / CUDA-event time.
cpu_baseline / --impl reference = the UNMODIFIED reference (oracle/_ref/orbit_ref, o=0) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

X_LOG2, Y_LOG2 = 24, 12          # BASELINE config[1]
METRIC = "orb_particle_passes_per_s"
UNIT = "particle-passes/s"


def workload_name(x_log2: int, y_log2: int, dist: str, n_levels: int, full: bool) -> str:
    """config.workload, shared by both arms so that the two JSON lines name the same workload"""
    return (f"2^{x_log2} {dist} particles per GPU (reference xorshf96 stream), 2^{y_log2} leaf cells, "
            f"{n_levels} split levels ({'full' if full else 'reference-compatible'})")


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------- reference arm (CPU)
def run_reference_cpu(x_log2: int, y_log2: int, threads: int, timeout: int = 1200):
    """Time the unmodified reference binary (oracle/_ref/orbit_ref, built by oracle/Makefile from the
    reference's own sources) in its CPU-only mode; falls back to the oracle port if the binary is absent."""
    ref_bin = ROOT / "oracle" / "_ref" / "orbit_ref"
    if ref_bin.exists():
        env = dict(os.environ, ORB_MDL_THREADS=str(threads))
        env.pop("ORB_REF_TRACE", None)
        r = subprocess.run([str(ref_bin), str(x_log2), str(y_log2), "0"], env=env, capture_output=True, text=True, timeout=timeout)
        if r.returncode == 0:
            out = r.stdout + r.stderr
            wall_us = int(re.search(r"RefBuildWall-us, (\d+)", out).group(1))
            passes = int(re.search(r"RefParticlePasses, (\d+)", out).group(1))
            lines = [l.strip() for l in r.stdout.splitlines() if l.strip()]
            return {"kind": "reference", "seconds": wall_us * 1e-6, "particle_passes": passes, "stdout": lines}
    # port: our C restatement, one shard per thread like the reference's mdl threads
    import numpy as np
    import oracle_py as oracle
    import orb_b200 as orb

    n = 1 << x_log2
    x, y, z = orb.generate_uniform(n)
    res = oracle.build(x, y, z, 1 << y_log2, ties=oracle.TIES_HOARE, n_shards=threads, n_threads=threads)
    st = res["stats"]
    return {"kind": "port", "seconds": st.t_total_s, "particle_passes": int(st.active_passes), "stdout": []}


def cpu_baseline_beside(x_log2: int, y_log2: int):
    """cpu_baseline of the GPU arm: the reference's CPU-only build of the same workload on all host cores, median of
    a few runs (a run is 0.15 s at 2^24 particles, 1.7 s at 2^27; bounded to ~10 s of CPU work)."""
    cores = host_cores()
    try:
        runs = [run_reference_cpu(x_log2, y_log2, cores)]
        budget_s = 10.0 - runs[0]["seconds"]
        while len(runs) < 5 and budget_s > runs[0]["seconds"]:
            runs.append(run_reference_cpu(x_log2, y_log2, cores))
            budget_s -= runs[-1]["seconds"]
        r = sorted(runs, key=lambda q: q["particle_passes"] / q["seconds"])[len(runs) // 2]
        return {"value": r["particle_passes"] / r["seconds"], "unit": UNIT, "cores": cores, "kind": r["kind"],
                "sample": f"full workload (orbit {x_log2} {y_log2} 0), median of {len(runs)} run(s), build wall {r['seconds']:.3f} s",
                "build_ms": r["seconds"] * 1e3}
    except Exception as exc:      # the baseline must never take the GPU numbers down with it
        return {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": f"failed: {exc}"}


def reference_arm(args):
    """--impl reference: the reference's own CPU path on the same config/metric, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    # weak scaling config: N GPUs <-> 2^24 * N particles; the CPU arm runs the per-GPU workload (a bounded
    # sample for N>1) - its throughput in particle-passes/s does not depend on which slice it builds
    times, passes, kind, lines = [], 0, "reference", []
    for i in range(args.warmup + args.steps):
        r = run_reference_cpu(X_LOG2, Y_LOG2, cores)
        kind, passes, lines = r["kind"], r["particle_passes"], r["stdout"]
        if i >= args.warmup:
            times.append(r["seconds"])
    sec = sum(times) / len(times)
    value = passes / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(X_LOG2, Y_LOG2, "uniform", Y_LOG2 - 1, False),
                   "particles_total": (1 << X_LOG2) * args.gpus, "leaf_cells": 1 << Y_LOG2,
                   "reference_run": f"orbit {X_LOG2} {Y_LOG2} 0 (CPU-only mode of the unmodified reference), one build of the per-GPU workload per step",
                   "levels": Y_LOG2 - 1, "threads": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"full workload (orbit {X_LOG2} {Y_LOG2} 0), {args.steps} run(s)", "reference_stdout": lines},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for l in self.proc.stdout:
            self.lines.append(l.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [t.strip() for t in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- our arm (GPU)
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--x", type=int, default=X_LOG2, help="log2 particles per GPU")
    ap.add_argument("--y", type=int, default=Y_LOG2, help="log2 leaf cells")
    ap.add_argument("--dist", default="uniform", choices=["uniform", "gaussian", "plummer"])
    ap.add_argument("--full-levels", action="store_true")
    ap.add_argument("--trial-depth", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    # stdout carries exactly one JSON line: whatever libraries print while the job runs (NCCL's version banner, ...)
    # goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import orb_b200 as orb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the ORB hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n_local, d = 1 << args.x, 1 << args.y
    W = max(args.warmup, 3)
    K = args.steps

    # ---- synthetic input: rank r takes slice [r*n, (r+1)*n) of the single reference stream (init.cu:11-25,47-53)
    if args.dist == "uniform":
        hx, hy, hz = orb.generate_uniform(n_local, skip=rank * n_local)
    else:
        hx, hy, hz = orb.generate_clustered(n_local, args.dist, skip=rank * n_local)
    pin = [torch.empty(n_local, dtype=torch.float32).pin_memory() for _ in range(3)]
    for t, h in zip(pin, (hx, hy, hz)):
        t.numpy()[:] = h
    pristine = [t.cuda(non_blocking=False) for t in pin]        # resident copy restored before every step
    out_pin = [torch.empty(n_local, dtype=torch.float32).pin_memory() for _ in range(3)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    ctx = orb.Orb(n_local, d, device=local)
    if args.trial_depth:
        ctx.set_trial_depth(args.trial_depth)
    use_peers = os.environ.get("ORB_NO_PEER", "0") != "1"
    if world > 1:
        from gpu_load_balance_b200 import dist as orb_dist

        orb_dist.connect(ctx, rank, world, device="cuda", peers=use_peers)

    def restore():
        ctx.load_device(pristine[0].data_ptr(), pristine[1].data_ptr(), pristine[2].data_ptr())
        flush.fill_(1)                       # L2 flush between timed iterations
        torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---- device-resident timing
    for _ in range(W):
        restore()
        ctx.build(full_levels=args.full_levels, want_heap=False)
    sampler = ClockSampler(local)
    if rank == 0:          # one nvidia-smi poller per job, not one per rank
        sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    ms_steps, launches, st_last = [], 0, None
    for _ in range(K):
        restore()
        barrier()
        _, st = ctx.build(full_levels=args.full_levels, want_heap=False)
        ms_steps.append(st.ms_total)
        launches += st.launches
        st_last = st
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    ms_total = float(sum(ms_steps))
    tt = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    pp = torch.tensor([float(st_last.iter_particle_passes)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)      # max over ranks
        dist.all_reduce(pp, op=dist.ReduceOp.SUM)      # whole-job particle passes
    ms_per_step = tt.item() / K
    passes_job = pp.item()
    value = passes_job / (ms_per_step * 1e-3)

    # ---- profiled steps: per-launch CUDA events on the library's stream (roofline)
    ctx.set_profile(True)
    prof = []
    for _ in range(3):
        restore()
        _, st = ctx.build(full_levels=args.full_levels, want_heap=False)
        prof.append(st)
    ctx.set_profile(False)
    stp = prof[-1]
    peak, peak_src = measured_peak_gbs()
    n_lv = stp.n_levels
    cnt_bytes = 4.0 * stp.active_passes                      # 4 B per active particle per HBM count pass
    part_bytes = 24.0 * n_local * n_lv                       # read+write x,y,z per level
    cnt_ms = sum(s.ms_count for s in prof) / len(prof)
    part_ms = sum(s.ms_partition for s in prof) / len(prof)
    cnt_gbs = cnt_bytes / (cnt_ms * 1e-3) / 1e9 if cnt_ms > 0 else 0.0
    part_gbs = part_bytes / (part_ms * 1e-3) / 1e9 if part_ms > 0 else 0.0
    tot_prof_ms = sum(s.ms_total for s in prof) / len(prof)
    kernels = {
        "k_count": {"bound": "hbm", "achieved": cnt_gbs, "peak": peak, "unit": "GB/s", "frac": cnt_gbs / peak, "traffic": None,
                    "launches_per_step": int(stp.count_launches), "ms_per_step": cnt_ms, "share_of_step": cnt_ms / tot_prof_ms,
                    "algorithmic_bytes_per_step": cnt_bytes},
        "k_partition": {"bound": "hbm", "achieved": part_gbs, "peak": peak, "unit": "GB/s", "frac": part_gbs / peak, "traffic": None,
                        "launches_per_step": int(stp.partition_launches), "ms_per_step": part_ms, "share_of_step": part_ms / tot_prof_ms,
                        "algorithmic_bytes_per_step": part_bytes},
    }
    # DRAM traffic per launch from the committed ncu --set full capture (profiles/r01_ncu_traffic.json)
    try:
        tr = json.loads((ROOT / "profiles" / "r01_ncu_traffic.json").read_text())
        for kname, rec in tr.items():
            if kname in kernels:
                kernels[kname]["traffic"] = rec["dram_bytes_read"] + rec["dram_bytes_write"]
                kernels[kname]["traffic_note"] = (f"ncu dram bytes of one launch ({rec['launch']}); algorithmic bytes of that "
                                                  f"launch: {rec['algorithmic_bytes_same_launch']}")
    except Exception:
        pass
    dom = "k_count" if cnt_ms >= part_ms else "k_partition"
    roofline = dict(kernels[dom], kernel=dom, peak_source=peak_src)

    # ---- the same build in the reference-exact tie mode (Hoare emulation), for information ----
    hoare = None
    try:
        ctx.set_tie_mode("hoare")
        hms = []
        for i in range(4):
            restore()
            _, sth = ctx.build(full_levels=args.full_levels, want_heap=False)
            if i:
                hms.append(sth.ms_total)
        hoare = {"build_ms": reduce_max(sum(hms) / len(hms)),
                 "note": "ORB_TIES_HOARE: cells, ranges and particle order equal to the reference CPU path even with tie particles"}
    except Exception as exc:   # informational leg must not break the bench line
        hoare = {"build_ms": None, "note": f"failed: {exc}"}
    finally:
        ctx.set_tie_mode("canonical")

    # ---- end to end through the public API with host buffers (H2D + build + D2H inside the timed region)
    e2e_times = []
    h2d = 12 * n_local
    d2h = 12 * n_local + ctx.n_heap * 52 + ctx.n_heap * 8
    for i in range(2 + min(K, 5)):
        barrier()
        t0 = time.perf_counter()
        ctx.upload(pin[0].numpy(), pin[1].numpy(), pin[2].numpy())
        heap, st = ctx.build(full_levels=args.full_levels, want_heap=True)
        ctx.download(tuple(t.numpy() for t in out_pin))
        rng = ctx.ranges()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i >= 2:
            e2e_times.append(dt)
    te = torch.tensor([sum(e2e_times) / len(e2e_times)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = te.item()
    e2e = {"value": passes_job / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}

    # ---- CPU baseline beside it (rank 0, N=1 only): the unmodified reference on all host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.dist == "uniform":
        cpu = cpu_baseline_beside(args.x, args.y)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": workload_name(args.x, args.y, args.dist, n_lv, args.full_levels),
                "particles_total": n_local * world, "leaf_cells": d, "parallelism": (f"particle shards x{world}; per-cell counts combined " +
                                                    ("by the selection search's two exchanges per level (NCCL allreduce of histogram rows, all-gather of candidates)"
                                                     if (world > 1 and os.environ.get("ORB_SELECT_MR", "1") != "0") else
                                                     ("inside the count kernel over NVLink peer memory" if (world > 1 and use_peers) else "by NCCL allreduce"))),
                "l2": "pristine particles restored + 256 MiB buffer written between timed steps (L2 flush)",
                "trial_depth": args.trial_depth or 3,
            },
            "build_ms": ms_per_step,
            "particle_passes_per_build": passes_job,
            "hbm_passes_particles_per_build_rank0": int(st_last.active_passes),
            "levels": n_lv, "iters": list(st_last.iters[:n_lv]), "passes": list(st_last.passes[:n_lv]),
            "wall_s_timed_region": t_wall,
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches,
            "roofline": roofline,
            "kernels": kernels,
            "cpu_baseline": cpu,
            "tie_mode": "canonical (stable x<cut; equals the reference whenever no particle sits exactly on a cut)",
            "reference_exact_mode": hoare,
        }
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
