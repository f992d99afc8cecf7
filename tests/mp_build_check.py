"""Helper launched by torchrun from tests/test_gpu_multiproc.py: one process per GPU, NCCL id broadcast, peer table
over CUDA IPC, whole build on this rank's slice, result dumped for comparison with the sharded oracle."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


FNV_OFFSET, FNV_PRIME = 1469598103934665603, 1099511628211


def known_answer_summary(out, rank, heap, st, rng, before, after):
    """Full-size runs (too large to ship particle dumps): digests comparable with `oracle/orb_oracle`'s stderr
    (iters, rangeHash of shard 0, heapHash) plus size-independent properties of this rank's slice."""
    import json
    import oracle_py

    L = st.n_levels
    ids = np.arange((1 << L) - 1, (1 << (L + 1)) - 1)
    b, e = rng[ids, 0].astype(np.int64), rng[ids, 1].astype(np.int64)
    n = before[0].size
    props = {"tiles": bool(b[0] == 0 and e[-1] == n and np.array_equal(b[1:], e[:-1]) and np.all(e >= b))}
    ne = e > b
    inside = True
    for a, col in enumerate(after):
        lo = np.minimum.reduceat(col, np.minimum(b, n - 1))[ne]
        hi = np.maximum.reduceat(col, np.minimum(b, n - 1))[ne]
        inside &= bool(np.all(lo >= heap["lower"][ids, a][ne]) and np.all(hi <= heap["upper"][ids, a][ne]))
    props["leaf_particles_inside_leaf_boxes"] = inside
    props["multiset_preserved"] = oracle_py.range_hashes(*before, 0, n)[0] == oracle_py.range_hashes(*after, 0, n)[0]
    rec = {"rank": rank, "props": props}
    import digests
    rec["digests"] = digests.rank_digests(rng, L, *after)
    rec["heapDigest"] = digests.heap_hash(heap)
    rec["iters_all"] = list(st.iters[:L])
    rec["not_found_all"] = list(st.not_found[:L])
    if rank == 0:
        h = FNV_OFFSET
        for v in e.tolist():
            h = ((h ^ v) * FNV_PRIME) & 0xFFFFFFFFFFFFFFFF
        hb = np.frombuffer(heap.tobytes(), np.uint8)
        rec.update(iters=list(st.iters[:L]), passes=list(st.passes[:L]), not_found=list(st.not_found[:L]),
                   rangeHash=f"{h:016x}", heapHash=f"{oracle_py.fnv1a(hb):016x}", ms_total=st.ms_total)
    (out / f"known{rank}.json").write_text(json.dumps(rec))


def tie_columns(n, seed=17):
    """A fifth of the particles share one value near the median on every axis (same recipe as
    tests/test_gpu_parity.py::test_search_fallback_on_massive_ties): the selection search has to flag cells."""
    rng = np.random.default_rng(seed)
    cols = []
    for a in range(3):
        v = (rng.random(n, dtype=np.float32) - 0.5).astype(np.float32)
        tie = rng.random(n) < 0.2
        v[tie] = np.float32(0.01 * (a + 1))
        cols.append(v)
    return cols


def main():
    import torch
    import torch.distributed as dist
    import orb_b200 as orb
    from gpu_load_balance_b200 import dist as od

    out = Path(sys.argv[1]); x_log2 = int(sys.argv[2]); y_log2 = int(sys.argv[3]); peers = sys.argv[4] == "1"
    known = len(sys.argv) > 5 and sys.argv[5] == "known"
    rank, world, local = od.env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_local, d = (1 << x_log2) // world, 1 << y_log2
    lo, _ = od.shard_slice(rank, world, n_local)
    ties = len(sys.argv) > 5 and sys.argv[5] == "ties"
    if ties:
        x, y, z = (col[lo:lo + n_local].copy() for col in tie_columns(n_local * world))
    else:
        x, y, z = orb.generate_uniform(n_local, skip=lo)
    ctx = orb.Orb(n_local, d, device=local)
    od.connect(ctx, rank, world, device="cuda", peers=peers)
    ctx.upload(x, y, z)
    heap, st = ctx.build()
    gx, gy, gz = ctx.download()
    rng = ctx.ranges()
    if known:
        known_answer_summary(out, rank, heap, st, rng, (x, y, z), (gx, gy, gz))
    else:
            np.savez(out / f"rank{rank}.npz", heap=heap.view(np.uint8), rng=rng, x=gx, y=gy, z=gz, iters=np.array(st.iters[:st.n_levels]),
                 not_found=np.array(st.not_found[:st.n_levels]), fallback=np.array([st.search_fallback_cells]),
                 passes=np.array(st.passes[:st.n_levels]))
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
