"""GPU test (needs >= 2 GPUs): one process per GPU under torchrun — NCCL communicator from a broadcast id, fused
count+combine over CUDA-IPC peer memory (and the NCCL-only path) — must reproduce the sharded oracle exactly."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("peers", [1, 0])
def test_two_process_build_matches_sharded_oracle(oracle, tmp_path, peers):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    x, y, R = 18, 8, 2
    port = 29600 + (os.getpid() % 300) + peers
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={R}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "tests" / "mp_build_check.py"), str(tmp_path), str(x), str(y), str(peers)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    xs, ys, zs = oracle.generate_uniform(1 << x)
    ref = oracle.build(xs, ys, zs, 1 << y, ties=oracle.TIES_CANONICAL, n_shards=R)
    per = (1 << x) // R
    for rank in range(R):
        got = np.load(tmp_path / f"rank{rank}.npz")
        assert got["heap"].tobytes() == ref["heap"].tobytes()
        assert np.array_equal(got["rng"], ref["ranges"][rank])
        assert np.array_equal(got["x"].view(np.uint32), ref["x"][rank * per:(rank + 1) * per].view(np.uint32))
        assert list(got["iters"]) == list(ref["stats"].iters[:len(got["iters"])])
