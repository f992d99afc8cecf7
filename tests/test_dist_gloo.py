"""CPU test of the N>1 host logic with torch.distributed/gloo, world_size 2: slice ownership, id
broadcast, and that per-rank counts combined by allreduce(sum) reproduce the single-rank bisection
(the only data-path exchange the ORB path has: Combine in countLeft.cpp:44-53)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, n_local, d, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import orb_b200 as orb
    import oracle_py as oracle
    from gpu_load_balance_b200 import dist as od

    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = od.shard_slice(rank, world, n_local)
    x, y, z = orb.generate_uniform(n_local, skip=lo)
    payload = bytes(range(128)) if rank == 0 else None
    got = od.broadcast_bytes(payload, 128, 0)
    assert got == bytes(range(128))
    # one level of the reference's loop, sharded: local count-left + allreduce(sum) + master decision
    cells = orb.root_cell(d)
    total = od.reduce_scalar(n_local, "sum")
    iters = 0
    while not cells[0]["foundCut"] and iters < 32:
        iters += 1
        cut = oracle.get_cut(cells[0])
        local = oracle.count_left(x, 0, n_local, cut)
        t = torch.tensor([local], dtype=torch.int64)
        dist.all_reduce(t)
        c = cells[0:1].copy()
        oracle.lib().orb_oracle_bisect_step(c.ctypes.data, int(t.item()), int(total))
        cells = c
    tmax = od.reduce_scalar(float(rank + 1), "max")
    q.put((rank, iters, float(oracle.get_cut(cells[0])), bool(cells[0]["foundCut"]), tmax, x[:4].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_bisection_matches_single_rank(oracle, orb):
    import torch.multiprocessing as mp

    world, n_local, d = 2, 1 << 14, 16
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_local, d, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-rank oracle on the concatenated stream
    x, y, z = oracle.generate_uniform(world * n_local)
    ref = oracle.build(x, y, z, d, ties=oracle.TIES_CANONICAL)
    want_cut = float(oracle.get_cut(ref["heap"][0]))
    for rank, iters, cut, found, tmax, head in res:
        assert iters == ref["stats"].iters[0]
        assert cut == want_cut and found
        assert tmax == float(world)
        assert head == x[rank * n_local: rank * n_local + 4].tolist()
