"""CPU tests of the tipsy snapshot reader/writer of the C++ host (gpu-load-balance_b200/host/tipsy), the input
path the reference declares but does not ship (init.cu:5,54-59; CMakeLists.txt:74).  Checked against an
independent numpy restatement of the file layout in tests/oracle_py.py."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "gpu-load-balance_b200" / "host"
TOOL = HOST / "tipsy_tool"


@pytest.fixture(scope="module")
def tool():
    subprocess.run(["make", "-C", str(HOST), str(TOOL)], check=True, capture_output=True)
    return TOOL


def positions(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.random((n, 3), dtype=np.float32) - 0.5).astype(np.float32)


@pytest.mark.parametrize("standard,header_bytes", [(True, 32), (False, 32), (False, 28)])
def test_reader_matches_numpy_layout(oracle, tool, tmp_path, standard, header_bytes):
    gas, dark, star = positions(1000, 1), positions(70001, 2), positions(333, 3)
    path = tmp_path / "snap.std"
    oracle.write_tipsy(path, gas, dark, star, standard=standard, header_bytes=header_bytes, time=0.125)
    info = json.loads(subprocess.run([str(tool), "info", str(path)], check=True, capture_output=True, text=True).stdout)
    assert info == {"count": 71334, "nsph": 1000, "ndark": 70001, "nstar": 333, "time": 0.125,
                    "standard": standard, "header_bytes": header_bytes}
    want = np.concatenate([gas, dark, star])
    # whole file, and slices that straddle the gas/dark and dark/star boundaries and the 65536-record read chunk
    for first, n in [(0, 71334), (990, 20), (70990, 344), (500, 70000), (71334, 0)]:
        out = tmp_path / "cols.raw"
        subprocess.run([str(tool), "dump", str(path), str(first), str(n), str(out)], check=True, capture_output=True)
        cols = np.fromfile(out, np.float32).reshape(3, n)
        assert np.array_equal(cols.view(np.uint32), want[first:first + n].T.view(np.uint32))


@pytest.mark.parametrize("kind", ["std", "native"])
def test_writer_round_trip(oracle, tool, tmp_path, kind):
    n = 70000
    pos = positions(n, 7)
    pos[0] = [-0.0, np.float32(1e-42), 0.5]      # signed zero and a denormal survive as bit patterns
    raw = tmp_path / "in.raw"
    np.ascontiguousarray(pos.T).tofile(raw)
    path = tmp_path / f"out.{kind}"
    subprocess.run([str(tool), "write", str(raw), str(n), str(path), kind], check=True, capture_output=True)
    x, y, z = oracle.read_tipsy(path)
    assert np.array_equal(np.stack([x, y, z]).view(np.uint32), np.ascontiguousarray(pos.T).view(np.uint32))
    info = json.loads(subprocess.run([str(tool), "info", str(path)], check=True, capture_output=True, text=True).stdout)
    assert info["count"] == n and info["ndark"] == n and info["standard"] == (kind == "std")


def test_reader_rejects_bad_files(tool, tmp_path):
    bad = tmp_path / "bad"
    bad.write_bytes(b"\0" * 10)
    assert subprocess.run([str(tool), "info", str(bad)], capture_output=True).returncode == 1
    bad.write_bytes(b"\1" * 64)                                   # ndim is neither 1..3 nor its swap
    r = subprocess.run([str(tool), "info", str(bad)], capture_output=True, text=True)
    assert r.returncode == 1 and "not a tipsy file" in r.stderr
    hdr = np.array([0.0], "<f8").tobytes() + np.array([10, 3, 0, 10, 0], "<i4").tobytes() + b"\0" * 4
    bad.write_bytes(hdr + b"\0" * (36 * 9))                       # one record short
    r = subprocess.run([str(tool), "info", str(bad)], capture_output=True, text=True)
    assert r.returncode == 1 and "size does not match" in r.stderr
    assert subprocess.run([str(tool), "info", str(tmp_path / "missing")], capture_output=True).returncode == 1
