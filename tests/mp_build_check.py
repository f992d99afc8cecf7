"""Helper launched by torchrun from tests/test_gpu_multiproc.py: one process per GPU, NCCL id broadcast, peer table
over CUDA IPC, whole build on this rank's slice, result dumped for comparison with the sharded oracle."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import torch
    import torch.distributed as dist
    import orb_b200 as orb
    from gpu_load_balance_b200 import dist as od

    out = Path(sys.argv[1]); x_log2 = int(sys.argv[2]); y_log2 = int(sys.argv[3]); peers = sys.argv[4] == "1"
    rank, world, local = od.env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_local, d = (1 << x_log2) // world, 1 << y_log2
    lo, _ = od.shard_slice(rank, world, n_local)
    x, y, z = orb.generate_uniform(n_local, skip=lo)
    ctx = orb.Orb(n_local, d, device=local)
    od.connect(ctx, rank, world, device="cuda", peers=peers)
    ctx.upload(x, y, z)
    heap, st = ctx.build()
    gx, gy, gz = ctx.download()
    rng = ctx.ranges()
    np.savez(out / f"rank{rank}.npz", heap=heap.view(np.uint8), rng=rng, x=gx, y=gy, z=gz, iters=np.array(st.iters[:st.n_levels]))
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
