"""TEST INFRASTRUCTURE: ctypes access to the CPU oracle (oracle/liborb_oracle.so), the trace
parser shared by oracle and reference traces, and helpers to run oracle/_ref/orbit_ref.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
import struct
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
ORACLE_LIB = ORACLE_DIR / "liborb_oracle.so"
REF_BIN = ORACLE_DIR / "_ref" / "orbit_ref"
GOLDEN = ROOT / "tests" / "golden"

CELL_DTYPE = np.dtype(
    [
        ("id", "<i4"), ("nLeafCells", "<i4"), ("prevCutAxis", "<i4"), ("cutAxis", "<i4"),
        ("foundCut", "u1"), ("pad_", "u1", (3,)),
        ("cutMarginLeft", "<f4"), ("cutMarginRight", "<f4"),
        ("lower", "<f4", (3,)), ("upper", "<f4", (3,)),
    ]
)

TIES_HOARE, TIES_CANONICAL = 0, 1
SID_INIT, SID_COUNTLEFT, SID_PARTITION, SID_COUNT, REC_PARTICLES = 2, 8, 9, 11, 1000


class Params(C.Structure):
    _fields_ = [
        ("d", C.c_int32), ("full_levels", C.c_int32), ("ties", C.c_int32), ("n_shards", C.c_int32),
        ("n_threads", C.c_int32), ("max_iter", C.c_int32), ("tight_box", C.c_int32),
        ("trace_particles", C.c_int32), ("trace_path", C.c_char_p),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("n_levels", C.c_int32), ("iters", C.c_int32 * 64), ("not_found", C.c_int32 * 64),
        ("active_passes", C.c_uint64), ("tie_particles", C.c_uint64),
        ("t_count_s", C.c_double), ("t_partition_s", C.c_double), ("t_makeaxis_s", C.c_double), ("t_total_s", C.c_double),
    ]


def build_oracle():
    subprocess.run(["make", "-C", str(ORACLE_DIR), "oracle"], check=True, capture_output=True)
    if (Path("/root/reference/src/orbit.cpp")).exists():
        subprocess.run(["make", "-C", str(ORACLE_DIR), "ref", "refbig"], check=True, capture_output=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not ORACLE_LIB.exists():
            build_oracle()
        L = C.CDLL(str(ORACLE_LIB))
        P = C.c_void_p
        L.orb_oracle_build.argtypes = [C.POINTER(Params), P, P, P, P, P, P, C.POINTER(Stats)]
        L.orb_oracle_build.restype = C.c_int
        L.orb_oracle_count_left.argtypes = [P, C.c_int64, C.c_int64, C.c_float]
        L.orb_oracle_count_left.restype = C.c_uint32
        L.orb_oracle_partition_canonical.argtypes = [P, P, P, C.c_int64, C.c_int64, C.c_int, C.c_float]
        L.orb_oracle_partition_canonical.restype = C.c_int64
        L.orb_oracle_bisect_step.argtypes = [P, C.c_uint32, C.c_uint32]
        L.orb_oracle_bisect_step.restype = C.c_int
        L.orb_oracle_cell_get_cut.argtypes = [P]
        L.orb_oracle_cell_get_cut.restype = C.c_float
        L.orb_oracle_bbox.argtypes = [P, P, P, C.c_int64, C.c_int64, P]
        L.orb_oracle_bbox.restype = None
        L.orb_oracle_generate_uniform.argtypes = [P, P, P, P, C.c_uint64]
        L.orb_oracle_xorshf96_init.argtypes = [P]
        L.orb_oracle_range_hashes.argtypes = [P, P, P, C.c_int64, C.c_int64, P, P]
        _lib = L
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


def generate_uniform(n: int):
    """Oracle's own restatement of init.cu:11-25,47-53."""
    st = (C.c_uint64 * 3)()
    lib().orb_oracle_xorshf96_init(st)
    x, y, z = (np.empty(n, np.float32) for _ in range(3))
    lib().orb_oracle_generate_uniform(st, _p(x), _p(y), _p(z), n)
    return x, y, z


def get_cut(cell) -> np.float32:
    """cell.h:74-76 in numpy: float add, halve in double, round to float."""
    s = np.float32(np.float32(cell["cutMarginRight"]) + np.float32(cell["cutMarginLeft"]))
    return np.float32(np.float64(s) / 2.0)


def count_left(col: np.ndarray, begin: int, end: int, cut) -> int:
    return int(lib().orb_oracle_count_left(_p(col), begin, end, C.c_float(float(cut))))


def bbox(x, y, z, begin, end) -> np.ndarray:
    out = np.zeros(6, np.float32)
    lib().orb_oracle_bbox(_p(x), _p(y), _p(z), begin, end, _p(out))
    return out


def partition_canonical(x, y, z, begin, end, axis, cut) -> int:
    return int(lib().orb_oracle_partition_canonical(_p(x), _p(y), _p(z), begin, end, axis, C.c_float(float(cut))))


def fnv1a(buf: np.ndarray) -> int:
    buf = np.ascontiguousarray(buf).view(np.uint8)
    f = lib().orb_oracle_fnv1a
    f.restype = C.c_uint64
    f.argtypes = [C.c_void_p, C.c_size_t]
    return int(f(buf.ctypes.data, buf.size))


def range_hashes(x, y, z, begin, end):
    a, b = C.c_uint64(0), C.c_uint64(0)
    lib().orb_oracle_range_hashes(_p(x), _p(y), _p(z), begin, end, C.byref(a), C.byref(b))
    return a.value, b.value


def build(x, y, z, d, *, ties=TIES_CANONICAL, full_levels=False, n_shards=1, n_threads=1, tight_box=False,
          trace_path=None, trace_particles=False, shard_off=None, copy=True):
    """Run the oracle build in place on copies of x,y,z (copy=False: on the arrays themselves, which must be
    contiguous float32). Returns dict(heap, ranges, stats, x, y, z)."""
    if copy:
        x = np.array(x, dtype=np.float32, copy=True)
        y = np.array(y, dtype=np.float32, copy=True)
        z = np.array(z, dtype=np.float32, copy=True)
    else:
        assert all(a.dtype == np.float32 and a.flags.c_contiguous for a in (x, y, z))
    n = x.size
    if shard_off is None:
        per = n // n_shards
        shard_off = np.array([per * s for s in range(n_shards)] + [per * n_shards], dtype=np.uint64)
    else:
        shard_off = np.asarray(shard_off, dtype=np.uint64)
    p = Params(d=d, full_levels=int(full_levels), ties=ties, n_shards=n_shards, n_threads=n_threads, max_iter=32,
               tight_box=int(tight_box), trace_particles=int(trace_particles),
               trace_path=(str(trace_path).encode() if trace_path else None))
    n_heap = 2 * d - 1
    heap = np.zeros(n_heap, CELL_DTYPE)
    ranges = np.zeros((n_shards, n_heap, 2), np.uint32)
    st = Stats()
    rc = lib().orb_oracle_build(C.byref(p), _p(x), _p(y), _p(z), _p(shard_off), _p(heap), _p(ranges), C.byref(st))
    if rc != 0:
        raise RuntimeError(f"orb_oracle_build failed: {rc}")
    return {"heap": heap, "ranges": ranges, "stats": st, "x": x, "y": y, "z": z, "shard_off": shard_off}


# ----------------------------------------------------------------------------- traces
def read_trace(path):
    """Parse an ORBTRACE file (written by oracle/ref_shim/ref_tap.cpp or orb_oracle.c)."""
    path = Path(path)
    raw = gzip.open(path, "rb").read() if path.suffix == ".gz" else path.read_bytes()
    assert raw[:8] == b"ORBTRACE", "not a trace"
    ver, cb = struct.unpack_from("<II", raw, 8)
    assert ver == 1 and cb == 52
    off = 16
    recs = []
    while off < len(raw):
        kind, n, nb = struct.unpack_from("<IIQ", raw, off)
        off += 16
        recs.append((kind, n, raw[off:off + nb]))
        off += nb
    return recs


def trace_levels(recs):
    """Group a trace into per-level dicts: cells at Count time, counts, per-iteration count-left
    (cells+counts), the partition record (final cells, child ranges, child hashes), particle dumps."""
    levels = []
    init_particles = None
    cur = None
    for kind, n, p in recs:
        if kind == SID_INIT:
            continue
        if kind == REC_PARTICLES:
            m = struct.unpack_from("<I", p, 0)[0]
            arr = np.frombuffer(p, dtype="<f4", offset=4).reshape(3, m).copy()
            if cur is None:
                init_particles = arr
            else:
                cur["particles"] = arr
            continue
        if kind == SID_COUNT:
            cur = {"cells": np.frombuffer(p[: n * 52], CELL_DTYPE).copy(),
                   "counts": np.frombuffer(p[n * 52:], "<u4").copy(), "iters": []}
            levels.append(cur)
        elif kind in (SID_COUNTLEFT, 6, 7):
            cur["iters"].append((np.frombuffer(p[: n * 52], CELL_DTYPE).copy(), np.frombuffer(p[n * 52:], "<u4").copy()))
        elif kind in (SID_PARTITION, 10):
            cur["final_cells"] = np.frombuffer(p[: n * 52], CELL_DTYPE).copy()
            rest = np.frombuffer(p[n * 52:], np.uint8).reshape(n, 48)
            cur["child_ranges"] = rest[:, :16].copy().view("<u4").reshape(n, 2, 2)
            cur["child_hashes"] = rest[:, 16:].copy().view("<u8").reshape(n, 2, 2)   # [cell][child][set, ordered]
    return init_particles, levels


REF_BIN_BIG = ORACLE_DIR / "_ref" / "orbit_ref_big"


def big_stack():
    """preexec_fn for orbit_ref_big: its per-level stack arrays / allocas grow with the lifted MAX_CELLS"""
    import resource

    resource.setrlimit(resource.RLIMIT_STACK, (1 << 30, resource.RLIM_INFINITY))


def run_reference(x: int, y: int, o: int = 0, trace_path=None, trace_particles=False, threads=1, timeout=600, big=False):
    """Run the reference binary built by oracle/Makefile (big: the build with MAX_CELLS lifted); returns its stdout+stderr."""
    env = dict(os.environ)
    env["ORB_MDL_THREADS"] = str(threads)
    if trace_path:
        env["ORB_REF_TRACE"] = str(trace_path)
    else:
        env.pop("ORB_REF_TRACE", None)
    if trace_particles:
        env["ORB_REF_TRACE_PARTICLES"] = "1"
    r = subprocess.run([str(REF_BIN_BIG if big else REF_BIN), str(x), str(y), str(o)], env=env, capture_output=True, text=True,
                       timeout=timeout, preexec_fn=big_stack if big else None)
    if r.returncode != 0:
        raise RuntimeError(f"orbit_ref failed: {r.returncode}\n{r.stdout}\n{r.stderr}")
    return r.stdout + r.stderr


def particle_hash(x, y, z) -> np.ndarray:
    """numpy version of orb_oracle_particle_hash (per particle, uint64)."""
    def mix(v):
        v = v ^ (v >> np.uint64(30)); v = v * np.uint64(0xBF58476D1CE4E5B9)
        v = v ^ (v >> np.uint64(27)); v = v * np.uint64(0x94D049BB133111EB)
        return v ^ (v >> np.uint64(31))
    with np.errstate(over="ignore"):
        xb = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
        yb = np.ascontiguousarray(y, np.float32).view(np.uint32).astype(np.uint64)
        zb = np.ascontiguousarray(z, np.float32).view(np.uint32).astype(np.uint64)
        a = (xb << np.uint64(32)) | yb
        b = zb | np.uint64(0x9E3779B900000000)
        return mix(a ^ mix(b))


def set_hash(x, y, z) -> int:
    with np.errstate(over="ignore"):
        return int(particle_hash(x, y, z).sum(dtype=np.uint64))


# ---- tipsy snapshots (checker-side numpy restatement of the file layout; the product reader is
# gpu-load-balance_b200/host/tipsy/tipsy.cpp, standing in for the reference's missing src/tipsy, init.cu:54-59) ----
TIPSY_FLOATS = {"gas": 12, "dark": 9, "star": 11}


def write_tipsy(path, gas, dark, star, *, standard=True, header_bytes=32, time=0.25, seed=0):
    """gas / dark / star: (n_k, 3) float32 positions; the other fields get arbitrary non-zero filler."""
    e = ">" if standard else "<"
    rng = np.random.default_rng(seed)
    parts = [np.asarray(p, np.float32).reshape(-1, 3) for p in (gas, dark, star)]
    n = [len(p) for p in parts]
    hdr = np.array([time], e + "f8").tobytes() + np.array([sum(n), 3, n[0], n[1], n[2]], e + "i4").tobytes()
    hdr += b"\0" * (header_bytes - 28)
    with open(path, "wb") as f:
        f.write(hdr)
        for pos, kind in zip(parts, ("gas", "dark", "star")):
            rec = rng.random((len(pos), TIPSY_FLOATS[kind]), dtype=np.float32) + 1.0
            rec[:, 1:4] = pos
            f.write(rec.astype(e + "f4").tobytes())


def read_tipsy(path):
    """Returns (x, y, z) float32 columns of all bodies, file order (gas, dark, star)."""
    raw = Path(path).read_bytes()
    e = "<" if 1 <= int(np.frombuffer(raw, "<i4", 1, 12)[0]) <= 3 else ">"
    n_all, ndim, ns, nd, nst = (int(v) for v in np.frombuffer(raw, e + "i4", 5, 8))
    assert ndim == 3 and ns + nd + nst == n_all
    body = 4 * (12 * ns + 9 * nd + 11 * nst)
    off = len(raw) - body
    assert off in (28, 32)
    cols = []
    for cnt, fl in ((ns, 12), (nd, 9), (nst, 11)):
        rec = np.frombuffer(raw, e + "f4", cnt * fl, off).reshape(cnt, fl)
        cols.append(rec[:, 1:4].astype(np.float32))
        off += 4 * cnt * fl
    pos = np.concatenate(cols)
    return (np.ascontiguousarray(pos[:, 0]), np.ascontiguousarray(pos[:, 1]), np.ascontiguousarray(pos[:, 2]))
