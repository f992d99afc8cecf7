"""GPU tests of the C++ host driver `orbit <x> <y> <o>` (same CLI/stdout as the reference): every mode
must leave exactly the oracle's tree, ranges and particle order; multi-rank runs (thread per GPU + NCCL)
must give the same cells as one rank."""
import os
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
ORBIT = ROOT / "gpu-load-balance_b200" / "host" / "orbit"


def run_orbit(x, y, o, tmp_path, threads=1, env_extra=None):
    subprocess.run(["make", "-C", str(ORBIT.parent)], check=True, capture_output=True)
    env = dict(os.environ, ORB_MDL_THREADS=str(threads), ORB_DUMP=str(tmp_path / "dump"))
    env.update(env_extra or {})
    r = subprocess.run([str(ORBIT), str(x), str(y), str(o)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def read_dump(tmp_path, oracle, rank=0):
    heap = np.fromfile(tmp_path / "dump.heap", dtype=oracle.CELL_DTYPE)
    heap["pad_"] = 0      # the 3 bytes after `bool foundCut` are compiler padding in the host's struct Cell
    raw = (tmp_path / f"dump.{rank}").read_bytes()
    n_heap, n = np.frombuffer(raw, "<u4", 2)
    off = 8
    rng = np.frombuffer(raw, "<u4", n_heap * 2, off).reshape(n_heap, 2); off += n_heap * 8
    x = np.frombuffer(raw, "<f4", n, off); off += 4 * n
    y = np.frombuffer(raw, "<f4", n, off); off += 4 * n
    z = np.frombuffer(raw, "<f4", n, off)
    return heap, rng, x, y, z


@pytest.mark.parametrize("o", [0, 1, 2])
@pytest.mark.parametrize("x,y", [(16, 6), (18, 9)])
def test_orbit_modes_match_oracle(oracle, tmp_path, x, y, o):
    out = run_orbit(x, y, o, tmp_path)
    # the three reference stdout lines (orbit.cpp:284-286)
    assert re.search(rf"^CountCopy-{x}-{y}, \d+ $", out, re.M)
    assert re.search(rf"^Partition-{x}-{y}, \d+ $", out, re.M)
    assert re.search(rf"^MakeAxis-{x}-{y}, 0 $", out, re.M)
    heap, rng, gx, gy, gz = read_dump(tmp_path, oracle)
    xs, ys, zs = oracle.generate_uniform(1 << x)
    ref = oracle.build(xs, ys, zs, 1 << y, ties=oracle.TIES_CANONICAL)
    assert heap.tobytes() == ref["heap"].tobytes()
    assert np.array_equal(rng, ref["ranges"][0])
    assert np.array_equal(gx.view(np.uint32), ref["x"].view(np.uint32))
    assert np.array_equal(gy.view(np.uint32), ref["y"].view(np.uint32))
    assert np.array_equal(gz.view(np.uint32), ref["z"].view(np.uint32))


def test_orbit_full_levels_env(oracle, tmp_path):
    run_orbit(15, 5, 0, tmp_path, env_extra={"ORB_FULL_LEVELS": "1"})
    heap, rng, *_ = read_dump(tmp_path, oracle)
    xs, ys, zs = oracle.generate_uniform(1 << 15)
    ref = oracle.build(xs, ys, zs, 1 << 5, ties=oracle.TIES_CANONICAL, full_levels=True)
    assert heap.tobytes() == ref["heap"].tobytes()
    assert np.array_equal(rng, ref["ranges"][0])


@pytest.mark.parametrize("o", [0, 1, 2])
def test_orbit_two_ranks_match_sharded_oracle(oracle, tmp_path, o):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    x, y, R = 17, 7, 2
    run_orbit(x, y, o, tmp_path, threads=R)
    xs, ys, zs = oracle.generate_uniform(1 << x)
    ref = oracle.build(xs, ys, zs, 1 << y, ties=oracle.TIES_CANONICAL, n_shards=R)
    heap = np.fromfile(tmp_path / "dump.heap", dtype=oracle.CELL_DTYPE)
    heap["pad_"] = 0
    assert heap.tobytes() == ref["heap"].tobytes()
    per = (1 << x) // R
    for r in range(R):
        _, rng, gx, gy, gz = read_dump(tmp_path, oracle, r)
        assert np.array_equal(rng, ref["ranges"][r])
        assert np.array_equal(gx.view(np.uint32), ref["x"][r * per:(r + 1) * per].view(np.uint32))


def test_orbit_reference_exact_ties_env(oracle, tmp_path):
    """ORB_TIES=hoare: the C++ host leaves exactly the reference CPU path's result (verbatim-Hoare oracle)."""
    x, y = 17, 6
    run_orbit(x, y, 0, tmp_path, env_extra={"ORB_TIES": "hoare"})
    heap, rng, gx, gy, gz = read_dump(tmp_path, oracle)
    xs, ys, zs = oracle.generate_uniform(1 << x)
    ref = oracle.build(xs, ys, zs, 1 << y, ties=oracle.TIES_HOARE)
    assert heap.tobytes() == ref["heap"].tobytes()
    assert np.array_equal(rng, ref["ranges"][0])
    assert np.array_equal(gx.view(np.uint32), ref["x"].view(np.uint32))
    assert np.array_equal(gz.view(np.uint32), ref["z"].view(np.uint32))


@pytest.mark.parametrize("o", [0, 2])
def test_orbit_two_ranks_reference_exact(oracle, tmp_path, o):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    x, y, R = 17, 7, 2
    run_orbit(x, y, o, tmp_path, threads=R, env_extra={"ORB_TIES": "hoare"})
    xs, ys, zs = oracle.generate_uniform(1 << x)
    ref = oracle.build(xs, ys, zs, 1 << y, ties=oracle.TIES_HOARE, n_shards=R)
    heap = np.fromfile(tmp_path / "dump.heap", dtype=oracle.CELL_DTYPE)
    heap["pad_"] = 0
    assert heap.tobytes() == ref["heap"].tobytes()
    per = (1 << x) // R
    for r in range(R):
        _, rng, gx, gy, gz = read_dump(tmp_path, oracle, r)
        assert np.array_equal(rng, ref["ranges"][r])
        assert np.array_equal(gx.view(np.uint32), ref["x"][r * per:(r + 1) * per].view(np.uint32))


def _snapshot(oracle, tmp_path, n):
    """Clustered positions in a big-endian tipsy file holding gas, dark and star bodies."""
    rng = np.random.default_rng(11)
    centres = rng.uniform(-0.35, 0.35, (16, 3))
    pos = (centres[rng.integers(0, 16, n)] + rng.normal(0.0, 0.03, (n, 3))).clip(-0.5, 0.5).astype(np.float32)
    path = tmp_path / "snap.std"
    oracle.write_tipsy(path, pos[:n // 5], pos[n // 5:n - 777], pos[n - 777:], standard=True)
    return path, pos


@pytest.mark.parametrize("o", [0, 1])
def test_orbit_tipsy_snapshot_input(oracle, tmp_path, o):
    """ORB_TIPSY: positions come from a tipsy snapshot (init.cu:54-59); x = 0 takes every body, here a count that
    is not a power of two.  Tree, ranges and particle order equal the oracle's on the same positions."""
    n, y = 50001, 6
    path, pos = _snapshot(oracle, tmp_path, n)
    out = run_orbit(0, y, o, tmp_path, env_extra={"ORB_TIPSY": str(path)})
    assert re.search(rf"^CountCopy-0-{y}, \d+ $", out, re.M)
    heap, rng, gx, gy, gz = read_dump(tmp_path, oracle)
    ref = oracle.build(pos[:, 0], pos[:, 1], pos[:, 2], 1 << y, ties=oracle.TIES_CANONICAL)
    assert heap.tobytes() == ref["heap"].tobytes()
    assert np.array_equal(rng, ref["ranges"][0])
    assert np.array_equal(gx.view(np.uint32), ref["x"].view(np.uint32))
    assert np.array_equal(gy.view(np.uint32), ref["y"].view(np.uint32))
    assert np.array_equal(gz.view(np.uint32), ref["z"].view(np.uint32))


def test_orbit_tipsy_prefix_and_errors(oracle, tmp_path):
    """x > 0 takes the first 2^x bodies; a snapshot that is too small is refused."""
    path, pos = _snapshot(oracle, tmp_path, 50001)
    run_orbit(15, 5, 0, tmp_path, env_extra={"ORB_TIPSY": str(path)})
    heap, rng, gx, *_ = read_dump(tmp_path, oracle)
    m = 1 << 15
    ref = oracle.build(pos[:m, 0], pos[:m, 1], pos[:m, 2], 1 << 5, ties=oracle.TIES_CANONICAL)
    assert heap.tobytes() == ref["heap"].tobytes()
    assert np.array_equal(gx.view(np.uint32), ref["x"].view(np.uint32))
    env = dict(os.environ, ORB_MDL_THREADS="1", ORB_TIPSY=str(path))
    r = subprocess.run([str(ORBIT), "16", "5", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "fewer than" in r.stderr


def test_orbit_tipsy_two_ranks_uneven_slices(oracle, tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n, y, R = 50001, 6, 2
    path, pos = _snapshot(oracle, tmp_path, n)
    run_orbit(0, y, 0, tmp_path, threads=R, env_extra={"ORB_TIPSY": str(path)})
    off = [r * n // R for r in range(R + 1)]
    ref = oracle.build(pos[:, 0], pos[:, 1], pos[:, 2], 1 << y, ties=oracle.TIES_CANONICAL, n_shards=R, shard_off=off)
    heap = np.fromfile(tmp_path / "dump.heap", dtype=oracle.CELL_DTYPE)
    heap["pad_"] = 0
    assert heap.tobytes() == ref["heap"].tobytes()
    for r in range(R):
        _, rng, gx, gy, gz = read_dump(tmp_path, oracle, r)
        assert np.array_equal(rng, ref["ranges"][r])
        assert np.array_equal(gx.view(np.uint32), ref["x"][off[r]:off[r + 1]].view(np.uint32))
        assert np.array_equal(gz.view(np.uint32), ref["z"][off[r]:off[r + 1]].view(np.uint32))
