"""CPU tests of the N>1 logic with torch.distributed/gloo, world_size 2: slice ownership, id broadcast, that per-rank
counts combined by allreduce(sum) reproduce the single-rank bisection (the only data-path exchange the reference has:
Combine in countLeft.cpp:44-53), and the selection search's two exchanges per level (histogram rows summed, candidate
slots gathered) restated in numpy over real collectives."""
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


def _select_worker(rank, world, port, n_local, nleaf, q):
    """The multi-rank selection search of one cell (csrc/orb_select.cuh: k_selx_resolve / k_selmr_finish) in numpy:
    exchange 1 sums the histogram rows, exchange 2 gathers fixed-size candidate slots whose last word is the count."""
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import orb_b200 as orb
    import test_select_model as m
    from gpu_load_balance_b200 import dist as od

    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, _ = od.shard_slice(rank, world, n_local)
    x, _, _ = orb.generate_uniform(n_local, skip=lo)
    x = x.copy()
    x[rank::97] = np.float32(0.0173)                       # a few hundred ties near the cut, spread over both ranks
    L, R, nb1, slot_words = np.float32(-0.5), np.float32(0.5), 512, 4096
    total = int(od.reduce_scalar(n_local, "sum"))
    prod = m.make_prod(total, nleaf)
    lo1, s1 = m.bin_params(L, R, nb1)
    b = m.sel_bin(x, lo1, s1, nb1)
    row_l = np.bincount(b, minlength=nb1).astype(np.int64)
    row_g = torch.from_numpy(row_l.copy())
    dist.all_reduce(row_g)                                 # exchange 1: rows over ranks
    p1 = np.concatenate([[0], np.cumsum(row_g.numpy())])
    f1, l1 = m.ambiguous_range(p1, 0, prod)                # every rank resolves the same candidate bins
    base = int(p1[f1])
    own = x[(b >= f1) & (b <= l1)]
    assert own.size <= slot_words - 1
    slot = np.zeros(slot_words, np.float32)
    slot[:own.size] = own
    slot.view(np.uint32)[-1] = own.size                    # count word
    parts = [torch.empty(slot_words, dtype=torch.float32) for _ in range(world)]
    dist.all_gather(parts, torch.from_numpy(slot))         # exchange 2: candidate slots
    counts = [int(p.numpy().view(np.uint32)[-1]) for p in parts]
    cand = np.concatenate([p.numpy()[:k] for p, k in zip(parts, counts)])
    assert cand.size == int(p1[l1 + 1] - p1[f1])
    # replay of orbit.cpp:149-232 on exchanged data only
    it, found, nleft_g = 0, False, None
    while it < m.MAX_ITER:
        cut = m.mid_cut(L, R)
        c1 = int(m.sel_bin(cut, lo1, s1, nb1))
        dec = -1 if c1 < f1 else (1 if c1 > l1 else 0)
        it += 1
        if dec == 0:
            cnt = base + int(np.count_nonzero(cand < cut))
            dv = m.diff_of(cnt, prod)
            if abs(dv) < 3:
                found, nleft_g = True, cnt
                break
            dec = 1 if dv > 0 else -1
        if dec > 0:
            R = cut
        else:
            L = cut
    cutf = m.mid_cut(L, R)
    nleft_l = int(row_l[:f1].sum()) + int(np.count_nonzero(own < cutf))      # own rows below the candidate bins + own candidates
    assert nleft_l == int(np.count_nonzero(x < cutf))
    q.put((rank, it, found, np.float32(L).tobytes(), np.float32(R).tobytes(), nleft_g, nleft_l, x))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nleaf", [(2, 2), (2, 7), (3, 5)])
def test_selection_search_protocol_over_ranks(world, nleaf):
    """Every rank ends with the literal loop's margins / iterations / global count, and the local left counts add up."""
    import torch.multiprocessing as mp
    import test_select_model as m

    n_local = 1 << 15
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + nleaf + 10 * world
    procs = [ctx.Process(target=_select_worker, args=(r, world, port, n_local, nleaf, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=240) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    v = np.concatenate([r[7] for r in res])
    want = m.literal_bisection(v, np.float32(-0.5), np.float32(0.5), v.size, nleaf)
    for rank, it, found, Lb, Rb, nleft_g, nleft_l, _ in res:
        assert (Lb, Rb, it, found) == (np.float32(want[0]).tobytes(), np.float32(want[1]).tobytes(), want[2], want[3])
        if found:
            assert nleft_g == want[4]
    assert sum(r[6] for r in res) == want[4]
