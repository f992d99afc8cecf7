"""Under torchrun: every rank builds its slice of 2^x particles per rank into 2^y leaf cells a few times; prints the
median build time (max over ranks).  With ORB_PROFILE=1 ORB_DEBUG_SELECT=1 the library prints per-kernel event times.
usage: torchrun --nproc-per-node R tools/mr_build_once.py [x_per_rank] [y] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import orb_b200 as orb
from gpu_load_balance_b200 import dist as od

x_log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 24
y_log2 = int(sys.argv[2]) if len(sys.argv) > 2 else 12
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rank, world, local = od.env_rank_world()
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, d = 1 << x_log2, 1 << y_log2
x, y, z = orb.generate_uniform(n, skip=rank * n)
ctx = orb.Orb(n, d, device=local)
od.connect(ctx, rank, world, device="cuda", peers=os.environ.get("ORB_NO_PEER", "0") != "1")
ms = []
for rep in range(1 + reps):
    ctx.upload(x, y, z)
    dist.barrier()
    if rank == 0 and rep == reps:
        print("==== last build ====", file=sys.stderr, flush=True)
    heap, st = ctx.build()
    if rep:
        ms.append(st.ms_total)
t = torch.tensor([float(np.median(ms))], dtype=torch.float64, device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("ranks %d ms median (max over ranks) %.4f" % (world, t.item()), "passes", list(st.passes[:st.n_levels]), "fallback_cells", st.search_fallback_cells)
ctx.close()
dist.barrier()
dist.destroy_process_group()
