"""CPU test of the C++ host's runtime (gpu-load-balance_b200/host: mdl threads, ServiceSetAdd, TraverseCombinePST):
for 1..8 rank threads every service call reaches every rank exactly once and the replies fold correctly.  The
driver proper (`orbit`) needs a GPU and is covered by tests/test_gpu_orbit_cli.py."""
import os
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "gpu-load-balance_b200" / "host"


@pytest.fixture(scope="module")
def check_binary(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostrt") / "host_traverse_check"
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-pthread", f"-I{HOST}", f"-I{HOST / 'mdl'}", str(ROOT / "tests" / "host_traverse_check.cpp"),
           str(HOST / "mdl" / "mdl.cpp"), str(HOST / "services" / "TraversePST.cpp"), str(HOST / "services" / "setadd.cpp"), "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


@pytest.mark.parametrize("threads", [1, 2, 3, 5, 8])
def test_service_calls_reach_every_rank_once(check_binary, threads):
    r = subprocess.run([str(check_binary)], env=dict(os.environ, ORB_MDL_THREADS=str(threads)), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"host runtime ok: {threads} ranks" in r.stdout
