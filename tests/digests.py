"""Digests of an ORB build result, pure numpy (no oracle, no GPU): shared by tests/golden/make_known_answers.py (which
applies them to the CPU oracle's output), bench.py's parity check and the full-size GPU tests (which apply them to what
liborb_b200.so produced).  All arithmetic is uint64 with wrap-around."""
from __future__ import annotations

import json
from pathlib import Path

import numpy as np

FNV_OFFSET, FNV_PRIME = 1469598103934665603, 1099511628211
KNOWN_PATH = Path(__file__).resolve().parent / "golden" / "known_answers.json"


def heap_hash(heap: np.ndarray) -> str:
    """Digest of the Cell heap.  Not byte-serial FNV (too slow in python at 2^21 cells): the heap is viewed as 13
    uint32 words per cell and folded with a position-weighted sum of mixed words, order sensitive."""
    w = np.ascontiguousarray(heap).view(np.uint8).reshape(-1).view("<u4").astype(np.uint64)
    with np.errstate(over="ignore"):
        v = _mix(w + np.uint64(0x9E3779B97F4A7C15) * np.arange(1, w.size + 1, dtype=np.uint64))
        return f"{int(v.sum(dtype=np.uint64)):016x}"


def _mix(v):
    v = v ^ (v >> np.uint64(30)); v = v * np.uint64(0xBF58476D1CE4E5B9)
    v = v ^ (v >> np.uint64(27)); v = v * np.uint64(0x94D049BB133111EB)
    return v ^ (v >> np.uint64(31))


def particle_hash(x, y, z) -> np.ndarray:
    """per particle uint64, equal to orb_oracle_particle_hash (oracle/orb_oracle.c) and tests/oracle_py.particle_hash"""
    with np.errstate(over="ignore"):
        xb = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
        yb = np.ascontiguousarray(y, np.float32).view(np.uint32).astype(np.uint64)
        zb = np.ascontiguousarray(z, np.float32).view(np.uint32).astype(np.uint64)
        a = (xb << np.uint64(32)) | yb
        b = zb | np.uint64(0x9E3779B900000000)
        return _mix(a ^ _mix(b))


def _fnv_fold(vals: np.ndarray) -> int:
    """h = OFFSET; for v: h = (h ^ v) * PRIME  (python ints; vals has one entry per leaf)"""
    h = FNV_OFFSET
    M = (1 << 64) - 1
    for v in vals.tolist():
        h = ((h ^ int(v)) * FNV_PRIME) & M
    return h


def leaf_ids(n_levels: int) -> np.ndarray:
    """heap ids of the leaves = children of the last split level (orbit.cpp:102: one level short of 2^y by default)"""
    return np.arange((1 << n_levels) - 1, (1 << (n_levels + 1)) - 1)


def rank_digests(ranges: np.ndarray, n_levels: int, x, y, z, chunk: int = 1 << 24) -> dict:
    """ranges: [nHeap][2] of one rank (cellToRangeMap); x, y, z: that rank's particles after the build."""
    ids = leaf_ids(n_levels)
    b = ranges[ids, 0].astype(np.int64)
    e = ranges[ids, 1].astype(np.int64)
    n = int(np.asarray(x).size)
    out = {"rangeHash": f"{_fnv_fold(e):016x}"}
    tiles = bool(n == 0 or (b[0] == 0 and e[-1] == n and np.array_equal(b[1:], e[:-1]) and np.all(e >= b)))
    out["leaves_tile_slice"] = tiles
    # per-particle hashes in chunks (2^27 particles x several uint64 temporaries would not fit comfortably otherwise)
    order = np.uint64(0)
    csum = np.zeros(n + 1, np.uint64)       # prefix sums of the particle hashes: leaf sums by difference
    with np.errstate(over="ignore"):
        for s in range(0, n, chunk):
            t = min(n, s + chunk)
            ph = particle_hash(x[s:t], y[s:t], z[s:t])
            idx = np.arange(s, t, dtype=np.uint64)
            order = order + (ph * (idx * np.uint64(2) + np.uint64(1))).sum(dtype=np.uint64)
            np.cumsum(ph, dtype=np.uint64, out=csum[s + 1:t + 1])
            if s:
                csum[s + 1:t + 1] += csum[s]
        leaf = csum[np.clip(e, 0, n)] - csum[np.clip(b, 0, n)] if tiles else np.zeros(len(ids), np.uint64)
    out["leafSetHash"] = f"{_fnv_fold(leaf):016x}"
    out["orderHash"] = f"{int(order):016x}"
    return out


def load_known() -> dict:
    return json.loads(KNOWN_PATH.read_text()) if KNOWN_PATH.exists() else {}


def known_key(config: str, n_ranks: int, ties: str = "canonical") -> str:
    return f"{config}_r{n_ranks}" + ("" if ties == "canonical" else "_hoare")


def compare(rec: dict, rank: int, *, iters, not_found, heap, ranges, n_levels, x, y, z, check_order: bool = True) -> dict:
    """Compare one rank's build result with a known-answers record; returns {"ok": bool, "mismatch": [names]}."""
    bad = []
    if list(iters) != rec["iters"]:
        bad.append("iters")
    if list(not_found) != rec["not_found"]:
        bad.append("not_found")
    if heap is not None and heap_hash(heap) != rec["heapHash"]:
        bad.append("heapHash")
    mine = rank_digests(ranges, n_levels, x, y, z)
    want = rec["ranks"][rank]
    for k in ("rangeHash", "leafSetHash") + (("orderHash",) if check_order else ()):
        if mine[k] != want[k]:
            bad.append(k)
    if not mine["leaves_tile_slice"]:
        bad.append("leaves_tile_slice")
    return {"ok": not bad, "mismatch": bad, "digests": mine}
