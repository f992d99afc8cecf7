#!/usr/bin/env python
"""bench.py — ORB build benchmark (BASELINE.json metric: build time & particle-passes/s vs HBM roofline).

A "step" is one full ORB build (cut search per level, split, partition) of a workload through the C ABI
(liborb_b200.so).  Workloads are BASELINE.json's configs, named by --config:

    c2   2^24 uniform particles -> 2^12 leaf cells           c4g / c4p   2^26 Gaussian / Plummer clumps -> 2^14
    c3   2^27 uniform particles -> 2^16 leaf cells           c5          2^30 uniform -> 2^20 (the north-star target)
    c2w  2^24 uniform particles PER GPU -> 2^12 (weak scaling; round 1's multi-GPU workload)

The particle count is the WHOLE job's: with N GPUs every rank holds the contiguous slice [r, r+1) * 2^x / N of the one
reference generator stream (the reference's static per-thread shards, orbit.cpp:83), so N = 1, 2, 4, 8 build the same
tree (strong scaling; digests are checked against the CPU oracle's for the same sharding).

Default run: N = 1 -> c3 (the largest single-GPU config of BASELINE.json) with c2, c4g, c4p as extra legs of the same
JSON line; N = 2, 4 -> c3 sharded; N = 8 -> c3 sharded with c5 as an extra leg.

value      = reference-equivalent particle-passes per second, whole job, particles resident in HBM: sum over bisection
             iterations of particles in still-unfound cells (what orbit.cpp's loop streams; identical for the reference
             and for this build because the cut sequence is bit-identical) / build time.
ms_per_step= ORB build time (CUDA events on the library's stream, max over ranks).
e2e        = same metric through the public API with HOST buffers: pinned host x,y,z -> device, build, device -> host
             x,y,z + cell heap + ranges, inside the timed region.
roofline   = dominant kernel group (largest share of the step), algorithmic bytes / CUDA-event time.
parity     = every leg's result (iterations, cell heap, leaf ranges, per-leaf particle sets, particle order) compared
             with tests/golden/known_answers.json (CPU oracle); a mismatch makes the exit code non-zero.
cpu_baseline / --impl reference = the reference's CPU path on all host cores: the UNMODIFIED reference binary
             (oracle/_ref/orbit_ref) where its MAX_CELLS = 8096 cap allows (y <= 12), else our cap-free C restatement.
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

METRIC = "orb_particle_passes_per_s"
UNIT = "particle-passes/s"

# x = log2 of ALL particles of the job, y = log2 leaf cells
CONFIGS = {
    "c1": dict(x=20, y=10, dist="uniform", baseline="configs[0]"),
    "c2": dict(x=24, y=12, dist="uniform", baseline="configs[1]"),
    "c3": dict(x=27, y=16, dist="uniform", baseline="configs[2]"),
    "c4g": dict(x=26, y=14, dist="gaussian", baseline="configs[3] (Gaussian clumps)"),
    "c4p": dict(x=26, y=14, dist="plummer", baseline="configs[3] (Plummer spheres)"),
    "c5": dict(x=30, y=20, dist="uniform", baseline="configs[4]"),
    "c2w": dict(x=24, y=12, dist="uniform", baseline="configs[1] per GPU (weak scaling)", weak=True),
}


def config_of(name: str, n_gpus: int) -> dict:
    c = dict(CONFIGS[name])
    if c.get("weak"):
        c["x"] += max(0, n_gpus.bit_length() - 1)
    c["name"] = name
    c["scaling"] = "weak" if c.get("weak") else "strong"
    return c


def default_legs(n_gpus: int):
    if n_gpus == 1:
        return ["c3", "c2", "c4g", "c4p"]
    if n_gpus >= 8:
        return ["c3", "c5"]
    return ["c3"]


def config_dict(cfg: dict, n_gpus: int, full: bool = False) -> dict:
    """`config` of the JSON line - the same dict in both arms, so that the two lines name the same workload."""
    levels = cfg["y"] if full else cfg["y"] - 1
    return {
        "workload": (f"{cfg['name']}: 2^{cfg['x']} {cfg['dist']} particles in all (reference xorshf96 stream), 2^{cfg['y']} leaf cells, "
                     f"{levels} split levels ({'full' if full else 'reference-compatible'}); BASELINE.json {cfg['baseline']}"),
        "particles_total": 1 << cfg["x"], "leaf_cells": 1 << cfg["y"], "levels": levels,
        "sharding": f"{n_gpus} contiguous slice(s) of the one generator stream, one per GPU / reference thread group",
        "l2": "GPU arm: pristine particles restored + 256 MiB buffer written between timed steps (L2 flush); inputs >= 192 MB",
    }


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------- reference arm (CPU)
_PARTICLES = {}


def run_reference_cpu(cfg: dict, threads: int, timeout: int = 1800):
    """One build of the workload by the reference's CPU path (o=0).  Uniform inputs of up to 2^12 leaf cells: the unmodified
    reference binary (oracle/_ref/orbit_ref, the reference's own sources compiled by oracle/Makefile).  More leaf cells: its
    MAX_CELLS = 8096 arrays overflow (constants.h:11), so oracle/_ref/orbit_ref_big runs - the same sources with that one
    constant lifted by the build recipe.  Clustered inputs (the reference has no such generator) or no binary: our C
    restatement of the same algorithm (oracle/orb_oracle.c), one shard per thread like the reference's mdl threads."""
    big = cfg["y"] > 12
    ref_bin = ROOT / "oracle" / "_ref" / ("orbit_ref_big" if big else "orbit_ref")
    if ref_bin.exists() and cfg["dist"] == "uniform":
        env = dict(os.environ, ORB_MDL_THREADS=str(threads))
        env.pop("ORB_REF_TRACE", None)

        def stack():     # the lifted MAX_CELLS grows the reference's stack arrays and allocas (TraversePST.cpp:41)
            import resource
            resource.setrlimit(resource.RLIMIT_STACK, (1 << 30, resource.RLIM_INFINITY))

        r = subprocess.run([str(ref_bin), str(cfg["x"]), str(cfg["y"]), "0"], env=env, capture_output=True, text=True, timeout=timeout,
                           preexec_fn=stack if big else None)
        if r.returncode == 0:
            out = r.stdout + r.stderr
            wall_us = int(re.search(r"RefBuildWall-us, (\d+)", out).group(1))
            passes = int(re.search(r"RefParticlePasses, (\d+)", out).group(1))
            lines = [l.strip() for l in r.stdout.splitlines() if l.strip()]
            return {"kind": "reference", "seconds": wall_us * 1e-6, "particle_passes": passes, "stdout": lines,
                    "binary": ref_bin.name + (" (reference sources, MAX_CELLS lifted by oracle/Makefile)" if big else " (unmodified reference)")}
    import oracle_py as oracle
    import orb_b200 as orb

    key = (cfg["x"], cfg["dist"])
    if key not in _PARTICLES:        # generated once per process; every run builds on a fresh copy
        _PARTICLES.clear()
        n = 1 << cfg["x"]
        _PARTICLES[key] = orb.generate_uniform(n) if cfg["dist"] == "uniform" else orb.generate_clustered(n, cfg["dist"])
    x, y, z = _PARTICLES[key]
    res = oracle.build(x, y, z, 1 << cfg["y"], ties=oracle.TIES_HOARE, n_shards=threads, n_threads=threads)
    st = res["stats"]
    return {"kind": "port", "seconds": st.t_total_s, "particle_passes": int(st.active_passes), "stdout": [],
            "binary": "oracle/liborb_oracle.so (C restatement; the reference has no clustered generator)"}


def cpu_baseline_beside(cfg: dict, budget_s: float = 12.0):
    """cpu_baseline of the GPU arm: the reference's CPU-only build of the same workload on all host cores, median of a
    few runs bounded to about `budget_s` of CPU work."""
    cores = host_cores()
    try:
        runs = [run_reference_cpu(cfg, cores)]
        left = budget_s - runs[0]["seconds"]
        while len(runs) < 5 and left > runs[0]["seconds"]:
            runs.append(run_reference_cpu(cfg, cores))
            left -= runs[-1]["seconds"]
        r = sorted(runs, key=lambda q: q["particle_passes"] / q["seconds"])[len(runs) // 2]
        return {"value": r["particle_passes"] / r["seconds"], "unit": UNIT, "cores": cores, "kind": r["kind"],
                "sample": f"full workload ({cfg['name']}: 2^{cfg['x']} -> 2^{cfg['y']}, CPU-only o=0 path), median of {len(runs)} run(s), "
                          f"build wall {r['seconds']:.3f} s",
                "build_ms": r["seconds"] * 1e3, "binary": r["binary"]}
    except Exception as exc:      # the baseline must never take the GPU numbers down with it
        return {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": f"failed: {exc}"}


def reference_arm(args):
    """--impl reference: the reference's own CPU path on the same config/metric, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    name = args.config or default_legs(args.gpus)[0]
    cfg = config_of(name, args.gpus)
    times, passes, kind, lines, binary = [], 0, "reference", [], ""
    t_start = time.time()
    for i in range(args.warmup + args.steps):
        r = run_reference_cpu(cfg, cores)
        kind, passes, lines, binary = r["kind"], r["particle_passes"], r["stdout"], r["binary"]
        if i >= args.warmup:
            times.append(r["seconds"])
        if time.time() - t_start > 240 and len(times) >= 1:     # bounded: the whole arm ends within a few minutes
            break
    sec = sum(times) / len(times)
    value = passes / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(cfg, args.gpus, args.full_levels),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"one full build of the workload per step ({len(times)} timed), CPU-only o=0 path on {cores} threads: {binary}",
                         "reference_stdout": lines},
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
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for l in self.proc.stdout:
            self.lines.append(l.strip())

    def wait_samples(self, n: int, timeout_s: float = 3.0):
        """block until at least n samples have arrived (short legs would otherwise end before the first one)"""
        t0 = time.time()
        while self.proc and len(self.lines) < n and time.time() - t0 < timeout_s:
            time.sleep(0.01)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for l in self.lines:
            f = [t.strip() for t in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(cfg_name: str, kernel: str):
    """DRAM bytes per launch of the kernel group from the kept `ncu --set full` capture of THIS config
    (profiles/r02_ncu_traffic.json, made by tools/ncu_traffic.py from the .ncu-rep); None if that config was not captured."""
    try:
        rec = json.loads((ROOT / "profiles" / "r02_ncu_traffic.json").read_text())[cfg_name][kernel]
        return rec["dram_bytes_read"] + rec["dram_bytes_write"], rec.get("note")
    except Exception:
        return None, None


# ----------------------------------------------------------------------------- our arm (GPU)
class Job:
    """process-wide plumbing: torch.distributed (NCCL) for barriers / reductions, nothing on the data path"""

    def __init__(self):
        import torch

        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the ORB hot path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, v, op="max"):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device="cuda")
        if self.dist is not None:
            self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN, "sum": self.dist.ReduceOp.SUM}[op])
        return t.item()


def run_leg(job: Job, name: str, args, primary: bool):
    """One workload: device-resident timing, per-kernel profile, parity digests, end-to-end through host buffers,
    reference-exact tie mode, CPU baseline.  Returns the leg's record (rank 0; other ranks return None)."""
    import numpy as np
    import orb_b200 as orb
    import digests

    torch = job.torch
    rank, world, local = job.rank, job.world, job.local
    cfg = config_of(name, world)
    n_total, d = 1 << cfg["x"], 1 << cfg["y"]
    n_local = n_total // world
    W = max(args.warmup, 3)
    K = args.steps if primary else max(3, min(args.steps, 10))
    full = args.full_levels

    # ---- synthetic input: rank r takes slice [r*n, (r+1)*n) of the single stream (init.cu:11-25,47-53)
    if cfg["dist"] == "uniform":
        hx, hy, hz = orb.generate_uniform(n_local, skip=rank * n_local)
    else:
        hx, hy, hz = orb.generate_clustered(n_local, cfg["dist"], skip=rank * n_local)
    pin = [torch.empty(n_local, dtype=torch.float32).pin_memory() for _ in range(3)]
    for t, h in zip(pin, (hx, hy, hz)):
        t.numpy()[:] = h
    del hx, hy, hz
    pristine = [t.cuda(non_blocking=False) for t in pin]        # resident copy restored before every step
    out_pin = [torch.empty(n_local, dtype=torch.float32).pin_memory() for _ in range(3)]

    ctx = orb.Orb(n_local, d, device=local)
    if args.trial_depth:
        ctx.set_trial_depth(args.trial_depth)
    use_peers = os.environ.get("ORB_NO_PEER", "0") != "1"
    if world > 1:
        from gpu_load_balance_b200 import dist as orb_dist

        orb_dist.connect(ctx, rank, world, device="cuda", peers=use_peers)

    def restore():
        ctx.load_device(pristine[0].data_ptr(), pristine[1].data_ptr(), pristine[2].data_ptr())
        job.flush.fill_(1)                   # L2 flush between timed iterations
        torch.cuda.synchronize()

    def check_parity(ties: str, heap, st):
        """digests of the result now on the device (and `heap`) against the CPU oracle's known answers"""
        key = digests.known_key(name, world, ties)
        rec = digests.load_known().get(key) if not full else None
        gx, gy, gz = ctx.download(tuple(t.numpy() for t in out_pin))
        rng = ctx.ranges()
        L = st.n_levels
        if rec is None:
            mine = digests.rank_digests(rng, L, gx, gy, gz)
            ok_all = job.reduce(1.0 if mine["leaves_tile_slice"] else 0.0, "min") > 0.5
            return {"status": "unpinned" if ok_all else "fail", "key": key, "note": "no known answers for this workload; leaf ranges tile the slice"
                    if ok_all else "leaf ranges do not tile the slice", "digests_rank0": mine, "heapHash": digests.heap_hash(heap)}
        res = digests.compare(rec, rank, iters=st.iters[:L], not_found=st.not_found[:L], heap=heap, ranges=rng, n_levels=L, x=gx, y=gy, z=gz)
        names = ["iters", "not_found", "heapHash", "rangeHash", "leafSetHash", "orderHash", "leaves_tile_slice"]
        mask = sum(1 << names.index(m) for m in res["mismatch"])
        # union of the mismatches over ranks (bit mask through a max-reduce per bit)
        bad = [nm for i, nm in enumerate(names) if job.reduce(1.0 if (mask >> i) & 1 else 0.0, "max") > 0.5]
        return {"status": "pass" if not bad else "fail", "key": key, "mismatch": bad, "ranks_checked": world,
                "checked": ["iters", "not_found", "heapHash", "rangeHash(per rank)", "leafSetHash(per rank)", "orderHash(per rank)"],
                "digests_rank0": res["digests"], "heapHash": digests.heap_hash(heap)}

    # ---- device-resident timing
    for _ in range(W):
        restore()
        ctx.build(full_levels=full, want_heap=False)
    sampler = ClockSampler(local)
    if rank == 0:          # one nvidia-smi poller per job, not one per rank
        sampler.start()
        sampler.wait_samples(1)
    job.barrier()
    t_wall0 = time.perf_counter()
    ms_steps, launches, st_last = [], 0, None
    for _ in range(K):
        restore()
        job.barrier()
        _, st = ctx.build(full_levels=full, want_heap=False)
        ms_steps.append(st.ms_total)
        launches += st.launches
        st_last = st
    job.barrier()
    t_wall = time.perf_counter() - t_wall0
    if rank == 0:
        sampler.wait_samples(3)
    clocks = sampler.stop()
    ms_per_step = job.reduce(float(sum(ms_steps)), "max") / K
    passes_job = job.reduce(float(st_last.iter_particle_passes), "sum")
    value = passes_job / (ms_per_step * 1e-3)
    fallback_cells = int(st_last.search_fallback_cells)

    # ---- parity of the canonical build (the one just timed, rebuilt once with the heap copied out)
    restore()
    heap, st = ctx.build(full_levels=full, want_heap=True)
    parity = check_parity("canonical", heap, st)

    # ---- profiled steps: per-launch CUDA events on the library's stream (roofline)
    ctx.set_profile(True)
    prof = []
    for _ in range(3):
        restore()
        _, st = ctx.build(full_levels=full, want_heap=False)
        prof.append(st)
    ctx.set_profile(False)
    stp = prof[-1]
    peak, peak_src = measured_peak_gbs()
    n_lv = stp.n_levels
    cnt_bytes = 4.0 * stp.active_passes                      # 4 B per active particle per HBM pass of the search
    part_bytes = 24.0 * n_local * n_lv                       # read+write x,y,z per level
    cnt_ms = job.reduce(sum(s.ms_count for s in prof) / len(prof), "max")
    part_ms = job.reduce(sum(s.ms_partition for s in prof) / len(prof), "max")
    tot_prof_ms = job.reduce(sum(s.ms_total for s in prof) / len(prof), "max")
    cnt_gbs = cnt_bytes / (cnt_ms * 1e-3) / 1e9 if cnt_ms > 0 else 0.0
    part_gbs = part_bytes / (part_ms * 1e-3) / 1e9 if part_ms > 0 else 0.0
    kernels = {}
    for kname, gbs, ms_k, nbytes, nl in (("k_count", cnt_gbs, cnt_ms, cnt_bytes, stp.count_launches),
                                         ("k_partition", part_gbs, part_ms, part_bytes, stp.partition_launches)):
        traffic, note = ncu_traffic(name if world == 1 else f"{name}_r{world}", kname)
        per_launch = nbytes / max(int(nl), 1)
        kernels[kname] = {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                          "traffic": traffic, "traffic_note": note, "launches_per_step": int(nl), "ms_per_step": ms_k,
                          "share_of_step": ms_k / tot_prof_ms if tot_prof_ms > 0 else None,
                          "algorithmic_bytes_per_step": nbytes, "algorithmic_bytes_per_launch": per_launch,
                          "note": "per rank (rank 0's bytes over the slowest rank's kernel time)"}
    dom = "k_count" if cnt_ms >= part_ms else "k_partition"
    roofline = dict(kernels[dom], kernel=dom, peak_source=peak_src)
    roofline["whole_step"] = {"algorithmic_bytes": cnt_bytes + part_bytes, "achieved": (cnt_bytes + part_bytes) / (ms_per_step * 1e-3) / 1e9,
                              "frac": (cnt_bytes + part_bytes) / (ms_per_step * 1e-3) / 1e9 / peak}

    # ---- end to end through the public API with host buffers (H2D + build + D2H inside the timed region)
    def e2e_leg(n_timed: int):
        times = []
        for i in range(2 + n_timed):
            job.barrier()
            t0 = time.perf_counter()
            ctx.upload(pin[0].numpy(), pin[1].numpy(), pin[2].numpy())
            ctx.build(full_levels=full, want_heap=True)
            ctx.download(tuple(t.numpy() for t in out_pin))
            ctx.ranges()
            torch.cuda.synchronize()
            if i >= 2:
                times.append(time.perf_counter() - t0)
        return job.reduce(sum(times) / len(times), "max")

    e2e = None
    h2d = 12 * n_local
    d2h = 12 * n_local + ctx.n_heap * 52 + ctx.n_heap * 8
    if primary or name == "c2":
        e2e_s = e2e_leg(min(K, 5))
        # the bare copies of the same bytes on all ranks at once: the machine's floor for this leg
        floor = []
        for i in range(4):
            job.barrier()
            t0 = time.perf_counter()
            for p, dv in zip(pin, pristine):
                dv.copy_(p, non_blocking=True)
            for o, dv in zip(out_pin, pristine):
                o.copy_(dv, non_blocking=True)
            torch.cuda.synchronize()
            if i:
                floor.append(time.perf_counter() - t0)
        floor_s = job.reduce(sum(floor) / len(floor), "max")
        e2e = {"value": passes_job / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "copy_floor_ms": floor_s * 1e3, "copy_floor_note": "pinned H2D + D2H of the particle columns alone, all ranks at once, max over ranks",
               "bytes_note": f"per rank ({world} rank(s) copy at once: the job moves {world} x these bytes per step)"}

    # ---- the same build in the reference-exact tie mode (Hoare emulation)
    hoare = None
    if world == 1 and (primary or name == "c2"):
        try:
            ctx.set_tie_mode("hoare")
            hms = []
            for i in range(4):
                restore()
                heap_h, sth = ctx.build(full_levels=full, want_heap=(i == 3))
                if i:
                    hms.append(sth.ms_total)
            hp = check_parity("hoare", heap_h, sth)
            he2e = e2e_leg(3)
            hoare = {"build_ms": sum(hms) / len(hms), "e2e_ms": he2e * 1e3, "e2e_value": passes_job / he2e, "parity": hp,
                     "note": "ORB_TIES_HOARE: cells, ranges and particle order equal to the reference CPU path even with tie particles"}
        except Exception as exc:   # informational leg must not break the bench line
            hoare = {"build_ms": None, "note": f"failed: {exc}"}
        finally:
            ctx.set_tie_mode("canonical")

    # ---- CPU baseline beside it (rank 0, N=1): the reference's CPU path on all host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and (primary or name == "c2"):
        cpu = cpu_baseline_beside(cfg)

    n_heap = ctx.n_heap
    ctx.close()
    del pristine, pin, out_pin
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    return {
        "name": name, "config": config_dict(cfg, world, full), "scaling": cfg["scaling"],
        "value": value, "ms_per_step": ms_per_step, "build_ms": ms_per_step, "steps": K, "warmup": W,
        "particle_passes_per_build": passes_job, "hbm_passes_particles_per_build_rank0": int(st_last.active_passes),
        "levels": n_lv, "iters": list(st_last.iters[:n_lv]), "passes": list(st_last.passes[:n_lv]),
        "search_fallback_cells": fallback_cells, "wall_s_timed_region": t_wall, "clocks": clocks, "parity": parity,
        "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
        "reference_exact_mode": hoare, "n_heap": n_heap,
        "parallelism": (f"particle shards x{world}; cut search exchanges per level: histogram rows reduced and candidates gathered to owner ranks "
                        "inside the kernels over NVLink peer memory (no collective call)" if (world > 1 and use_peers) else
                        (f"particle shards x{world}; NCCL allreduce / all-gather (ORB_NO_PEER=1)" if world > 1 else "single GPU")),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS), help="primary workload (default: c3)")
    ap.add_argument("--legs", default=None, help="comma-separated extra workloads in the same line ('' for none)")
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

    job = Job()
    names = default_legs(job.world)
    if args.config:
        names = [args.config]
    if args.legs is not None:
        names = names[:1] + [n for n in args.legs.split(",") if n]
    legs = []
    for i, nm in enumerate(names):
        legs.append(run_leg(job, nm, args, primary=(i == 0)))
    rc = 0
    if job.rank == 0:
        p = legs[0]
        line = {
            "metric": METRIC, "value": p["value"], "unit": UNIT, "n_gpus": job.world, "steps": p["steps"], "warmup": p["warmup"],
            "ms_per_step": p["ms_per_step"], "higher_is_better": True, "scaling": p["scaling"], "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": p["config"],
        }
        for k in ("build_ms", "particle_passes_per_build", "hbm_passes_particles_per_build_rank0", "levels", "iters", "passes",
                  "search_fallback_cells", "wall_s_timed_region", "clocks", "parity", "e2e", "gpu_launches", "roofline", "kernels",
                  "cpu_baseline", "reference_exact_mode", "parallelism"):
            line[k] = p[k]
        line["tie_mode"] = "canonical (stable x<cut; equals the reference whenever no particle sits exactly on a cut); reference_exact_mode = Hoare-exact"
        line["legs"] = {l["name"]: {k: v for k, v in l.items() if k != "name"} for l in legs[1:]}
        bad = [l["name"] for l in legs if l["parity"]["status"] == "fail" or
               (l.get("reference_exact_mode") and (l["reference_exact_mode"].get("parity") or {}).get("status") == "fail")]
        line["parity_all_legs"] = "pass" if not bad else f"fail: {','.join(bad)}"
        if bad:
            rc = 3
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if job.dist is not None:
        job.dist.barrier()
        job.dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
