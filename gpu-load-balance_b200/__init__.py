"""gpu-load-balance_b200 — B200-native ORB hot path (count-left + bisection, partition, bbox).

Python mirror of the C ABI in ``include/orb_b200.h`` (ctypes, no torch types cross the
boundary).  The directory name contains a hyphen, so import it through the helper::

    import orb_b200            # repo root shim -> this package

Everything here calls ``liborb_b200.so`` (hand-written sm_100a kernels).  There is NO CPU
fallback: if the library is missing or no CUDA device is present the calls raise.

Names follow the reference's services (src/services/*.h of andrinr/gpu-load-balance):
``count`` = ServiceCount, ``count_left`` = ServiceCountLeft*/GPU, ``partition`` =
ServicePartition*/GPU; ``find_cuts`` and ``build`` are the fused device-side loops of
orbit.cpp:146-232 and orbit.cpp:74-275.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
LIB_PATH = PKG_DIR / "liborb_b200.so"
HEADER = REPO_ROOT / "include" / "orb_b200.h"

ORB_OK = 0
ORB_FULL_LEVELS = 1
ORB_TIGHT_BOX = 2

# layout of the reference's struct Cell (cell.h:9-17), 52 bytes
CELL_DTYPE = np.dtype(
    [
        ("id", "<i4"),
        ("nLeafCells", "<i4"),
        ("prevCutAxis", "<i4"),
        ("cutAxis", "<i4"),
        ("foundCut", "u1"),
        ("pad_", "u1", (3,)),
        ("cutMarginLeft", "<f4"),
        ("cutMarginRight", "<f4"),
        ("lower", "<f4", (3,)),
        ("upper", "<f4", (3,)),
    ]
)
assert CELL_DTYPE.itemsize == 52


class BuildStats(C.Structure):
    _fields_ = [
        ("n_levels", C.c_int32),
        ("iters", C.c_int32 * 64),
        ("passes", C.c_int32 * 64),
        ("not_found", C.c_int32 * 64),
        ("active_passes", C.c_uint64),
        ("iter_particle_passes", C.c_uint64),
        ("count_launches", C.c_uint64),
        ("update_launches", C.c_uint64),
        ("partition_launches", C.c_uint64),
        ("other_launches", C.c_uint64),
        ("ms_count", C.c_float),
        ("ms_partition", C.c_float),
        ("ms_total", C.c_float),
        ("search_fallback_cells", C.c_uint32),
    ]

    def as_dict(self):
        n = self.n_levels
        return {
            "n_levels": n,
            "iters": list(self.iters[:n]),
            "passes": list(self.passes[:n]),
            "not_found": list(self.not_found[:n]),
            "active_passes": int(self.active_passes),
            "iter_particle_passes": int(self.iter_particle_passes),
            "count_launches": int(self.count_launches),
            "update_launches": int(self.update_launches),
            "partition_launches": int(self.partition_launches),
            "other_launches": int(self.other_launches),
            "ms_count": float(self.ms_count),
            "ms_partition": float(self.ms_partition),
            "ms_total": float(self.ms_total),
        }

    @property
    def launches(self):
        return int(self.count_launches + self.update_launches + self.partition_launches + self.other_launches)


class PeerInfo(C.Structure):
    """orb_peer_info: descriptor of one rank's counter rows for the fused count+combine (include/orb_b200.h)."""

    _fields_ = [
        ("ipc_cnt", C.c_uint8 * 64),
        ("ipc_flag", C.c_uint8 * 64),
        ("ptr_cnt", C.c_uint64),
        ("ptr_flag", C.c_uint64),
        ("pid", C.c_int64),
        ("device", C.c_int32),
        ("reserved_", C.c_int32),
        ("ipc_xchg", C.c_uint8 * 64),
        ("ptr_xchg", C.c_uint64),
    ]


PEER_INFO_BYTES = C.sizeof(PeerInfo)


class LevelPlan(C.Structure):
    """orb_level_plan: how a level would be searched (include/orb_b200.h, orb_plan_level)."""

    _fields_ = [
        ("search", C.c_int32),
        ("hist_bins", C.c_int32),
        ("cand_cap", C.c_uint32),
        ("slot_words", C.c_uint32),
        ("hist_words", C.c_uint64),
        ("prefuse_bins", C.c_int32),
        ("sample_stride", C.c_int32),
    ]


class OrbError(RuntimeError):
    pass


def build_library(force: bool = False) -> Path:
    """Compile liborb_b200.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    srcs = sorted((PKG_DIR / "csrc").glob("*.cu*")) + sorted((PKG_DIR / "csrc").glob("*.cpp")) + [PKG_DIR / "csrc" / "Makefile", HEADER]
    if not force and LIB_PATH.exists() and all(LIB_PATH.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return LIB_PATH
    subprocess.run(["make", "-C", str(PKG_DIR / "csrc")] + (["-B"] if force else []), check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the C-ABI library; fail loudly if it is missing (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise OrbError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the ORB hot path has no CPU fallback)"
        )
    L = C.CDLL(str(LIB_PATH))
    P = C.c_void_p
    u32p = C.POINTER(C.c_uint32)
    f32p = C.POINTER(C.c_float)
    sig = {
        "orb_create": ([C.POINTER(P), C.c_int, C.c_uint64, C.c_uint32], C.c_int),
        "orb_destroy": ([P], C.c_int),
        "orb_last_error": ([], C.c_char_p),
        "orb_version": ([], C.c_int),
        "orb_set_trial_depth": ([P, C.c_int], C.c_int),
        "orb_set_profile": ([P, C.c_int], C.c_int),
        "orb_plan_level": ([C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(LevelPlan)], C.c_int),
        "orb_set_tie_mode": ([P, C.c_int], C.c_int),
        "orb_comm_unique_id": ([P], C.c_int),
        "orb_comm_init": ([P, P, C.c_int, C.c_int], C.c_int),
        "orb_comm_attach": ([P, P, C.c_int, C.c_int], C.c_int),
        "orb_peer_export": ([P, P], C.c_int),
        "orb_peer_import": ([P, P, C.c_int], C.c_int),
        "orb_upload_xyz": ([P, P, P, P], C.c_int),
        "orb_load_device_xyz": ([P, P, P, P], C.c_int),
        "orb_download_xyz": ([P, P, P, P], C.c_int),
        "orb_device_xyz": ([P, C.POINTER(P), C.POINTER(P), C.POINTER(P)], C.c_int),
        "orb_count": ([P, P, C.c_uint32, P], C.c_int),
        "orb_count_left": ([P, P, C.c_uint32, P], C.c_int),
        "orb_partition": ([P, P, C.c_uint32], C.c_int),
        "orb_bbox": ([P, P, C.c_uint32, P], C.c_int),
        "orb_get_ranges": ([P, C.c_uint32, C.c_uint32, P], C.c_int),
        "orb_find_cuts": ([P, P, C.c_uint32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)], C.c_int),
        "orb_build": ([P, C.c_uint32, P, C.POINTER(BuildStats)], C.c_int),
        "orb_generate_uniform": ([C.c_uint64, C.c_uint64, P, P, P], None),
        "orb_generate_clustered": ([C.c_int, C.c_uint64, C.c_uint64, P, P, P], None),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    _lib = L
    return L


def _check(rc: int, what: str):
    if rc != ORB_OK:
        raise OrbError(f"{what} failed ({rc}): {lib().orb_last_error().decode(errors='replace')}")


def _f32(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def plan_level(n_local: int, n_cells: int, n_leaf_cells: int, n_ranks: int = 1, n_global: int | None = None,
               n_local_min: int | None = None, prefuse: int = -1) -> LevelPlan:
    """orb_plan_level: pure host arithmetic, works without a GPU."""
    out = LevelPlan()
    _check(lib().orb_plan_level(n_local, n_global if n_global is not None else n_local * n_ranks,
                                n_local_min if n_local_min is not None else n_local, n_ranks, n_leaf_cells, n_cells, prefuse,
                                C.byref(out)), "orb_plan_level")
    return out


# ----------------------------------------------------------------------------- inputs
def generate_uniform(n: int, skip: int = 0):
    """Reference generator (init.cu:11-25,47-53): returns x, y, z float32 arrays of length n."""
    x = np.empty(n, np.float32)
    y = np.empty(n, np.float32)
    z = np.empty(n, np.float32)
    lib().orb_generate_uniform(skip, n, _ptr(x), _ptr(y), _ptr(z))
    return x, y, z


def generate_clustered(n: int, kind: str = "gaussian", skip: int = 0):
    """Clustered inputs of SURVEY.md §8(d): 'gaussian' clumps or 'plummer' spheres."""
    x = np.empty(n, np.float32)
    y = np.empty(n, np.float32)
    z = np.empty(n, np.float32)
    lib().orb_generate_clustered(0 if kind == "gaussian" else 1, skip, n, _ptr(x), _ptr(y), _ptr(z))
    return x, y, z


def root_cell(d: int) -> np.ndarray:
    """Root cell as master() builds it (orbit.cpp:45-46,74-76)."""
    c = np.zeros(1, CELL_DTYPE)
    c["id"] = 0
    c["nLeafCells"] = d
    c["prevCutAxis"] = -1
    c["cutAxis"] = 0
    c["lower"] = -0.5
    c["upper"] = 0.5
    c["cutMarginLeft"] = -0.5
    c["cutMarginRight"] = 0.5
    return c


# ----------------------------------------------------------------------------- context
class Orb:
    """One ORB context = one GPU rank (the reference's per-thread LocalData, pst.h:6-41)."""

    def __init__(self, n_local: int, n_leaf_cells: int, device: int = 0):
        self._h = C.c_void_p()
        self.n_local = int(n_local)
        self.d = int(n_leaf_cells)
        self.n_heap = 2 * self.d - 1
        _check(lib().orb_create(C.byref(self._h), device, self.n_local, self.d), "orb_create")

    def close(self):
        if self._h:
            lib().orb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- tuning
    def set_trial_depth(self, m: int):
        _check(lib().orb_set_trial_depth(self._h, m), "orb_set_trial_depth")

    def set_tie_mode(self, mode: str):
        """'canonical' (stable x<cut, default) or 'hoare' (reference-exact ties and particle order)."""
        _check(lib().orb_set_tie_mode(self._h, {"canonical": 0, "hoare": 1}[mode]), "orb_set_tie_mode")

    def set_profile(self, on: bool):
        _check(lib().orb_set_profile(self._h, int(on)), "orb_set_profile")

    # ---- multi-GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(lib().orb_comm_unique_id(buf), "orb_comm_unique_id")
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, n_ranks: int):
        assert len(unique_id) == 128
        buf = C.create_string_buffer(unique_id, 128)
        _check(lib().orb_comm_init(self._h, buf, rank, n_ranks), "orb_comm_init")

    def peer_export(self) -> bytes:
        """This rank's peer descriptor (PEER_INFO_BYTES bytes) for the fused count+combine."""
        info = PeerInfo()
        _check(lib().orb_peer_export(self._h, C.byref(info)), "orb_peer_export")
        return bytes(info)

    def peer_import(self, table: bytes, n_ranks: int):
        """Descriptors of all ranks, concatenated in rank order (gathered by the caller)."""
        assert len(table) == n_ranks * PEER_INFO_BYTES
        buf = C.create_string_buffer(table, len(table))
        _check(lib().orb_peer_import(self._h, buf, n_ranks), "orb_peer_import")

    # ---- particles
    def upload(self, x, y, z):
        x, y, z = _f32(x), _f32(y), _f32(z)
        assert x.size == y.size == z.size == self.n_local
        _check(lib().orb_upload_xyz(self._h, _ptr(x), _ptr(y), _ptr(z)), "orb_upload_xyz")

    def load_device(self, dx: int, dy: int, dz: int):
        """Device-to-device load from raw device pointers (e.g. torch tensors' data_ptr())."""
        _check(lib().orb_load_device_xyz(self._h, C.c_void_p(dx), C.c_void_p(dy), C.c_void_p(dz)), "orb_load_device_xyz")

    def download(self, out=None):
        if out is None:
            out = tuple(np.empty(self.n_local, np.float32) for _ in range(3))
        x, y, z = out
        _check(lib().orb_download_xyz(self._h, _ptr(x), _ptr(y), _ptr(z)), "orb_download_xyz")
        return x, y, z

    # ---- service-granular calls
    def count(self, cells: np.ndarray) -> np.ndarray:
        cells = np.ascontiguousarray(cells, dtype=CELL_DTYPE)
        out = np.zeros(cells.size, np.uint32)
        _check(lib().orb_count(self._h, _ptr(cells), cells.size, _ptr(out)), "orb_count")
        return out

    def count_left(self, cells: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        cells = np.ascontiguousarray(cells, dtype=CELL_DTYPE)
        if out is None:
            out = np.zeros(cells.size, np.uint32)
        _check(lib().orb_count_left(self._h, _ptr(cells), cells.size, _ptr(out)), "orb_count_left")
        return out

    def partition(self, cells: np.ndarray):
        cells = np.ascontiguousarray(cells, dtype=CELL_DTYPE)
        _check(lib().orb_partition(self._h, _ptr(cells), cells.size), "orb_partition")

    def bbox(self, cells: np.ndarray) -> np.ndarray:
        cells = np.ascontiguousarray(cells, dtype=CELL_DTYPE)
        out = np.zeros((cells.size, 6), np.float32)
        _check(lib().orb_bbox(self._h, _ptr(cells), cells.size, _ptr(out)), "orb_bbox")
        return out

    def ranges(self, first_id: int = 0, n: int | None = None) -> np.ndarray:
        if n is None:
            n = self.n_heap - first_id
        out = np.zeros((n, 2), np.uint32)
        _check(lib().orb_get_ranges(self._h, first_id, n, _ptr(out)), "orb_get_ranges")
        return out

    # ---- fused calls
    def find_cuts(self, cells: np.ndarray):
        """Whole bisection loop of a level on the device; returns (cells, iterations, passes)."""
        cells = np.ascontiguousarray(cells, dtype=CELL_DTYPE).copy()
        it, ps = C.c_int32(0), C.c_int32(0)
        _check(lib().orb_find_cuts(self._h, _ptr(cells), cells.size, C.byref(it), C.byref(ps)), "orb_find_cuts")
        return cells, it.value, ps.value

    def build(self, full_levels: bool = False, tight_box: bool = False, want_heap: bool = True):
        """Whole ORB build on the device; returns (heap cells or None, BuildStats)."""
        flags = (ORB_FULL_LEVELS if full_levels else 0) | (ORB_TIGHT_BOX if tight_box else 0)
        heap = np.zeros(self.n_heap, CELL_DTYPE) if want_heap else None
        st = BuildStats()
        _check(lib().orb_build(self._h, flags, _ptr(heap) if want_heap else None, C.byref(st)), "orb_build")
        return heap, st


# ----------------------------------------------------------------------------- header <-> library check
def declared_symbols() -> list[str]:
    """Function names declared in include/orb_b200.h (used by the CPU-side ABI test)."""
    import re

    txt = HEADER.read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(orb_[a-z0-9_]+)\s*\(", txt)))
