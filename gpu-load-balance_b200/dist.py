"""Host-side plumbing for the one-process-per-GPU launch (torchrun): which slice of the single
particle stream a rank owns, how the NCCL unique id reaches every rank, and device-timed max-over-ranks.
torch.distributed is plumbing only; the data-path collective (per-cell count allreduce) is NCCL inside
liborb_b200.so.  Works with the gloo backend on CPU (tests/test_dist_gloo.py) and nccl on GPUs."""
from __future__ import annotations

import os


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_slice(rank: int, world: int, n_local: int):
    """Rank r owns particles [r*n_local, (r+1)*n_local) of the single generator stream — the reference's
    static per-thread shards (orbit.cpp:83: N/Threads particles per thread, never migrated)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    return rank * n_local, (rank + 1) * n_local


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0, device="cpu") -> bytes:
    """Send `nbytes` raw bytes from rank `src` to every rank (used for the 128-byte NCCL unique id)."""
    import torch
    import torch.distributed as dist

    buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        assert payload is not None and len(payload) == nbytes
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(buf, src)
    return bytes(buf.cpu().numpy().tobytes())


def reduce_scalar(value: float, op: str = "max", device="cpu") -> float:
    """max / sum of a scalar over ranks (timings: max; work counters: sum)."""
    import torch
    import torch.distributed as dist

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def all_gather_bytes(payload: bytes, device="cpu") -> bytes:
    """Concatenation over ranks (in rank order) of equal-length byte strings (peer descriptors)."""
    import torch
    import torch.distributed as dist

    mine = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(device)
    parts = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, mine)
    return b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)


def connect(ctx, rank: int, world: int, device="cuda", peers: bool = True):
    """Give an Orb context its communicator (NCCL id broadcast) and, optionally, the peer-memory table that lets
    the count kernel combine counts over NVLink by itself."""
    from . import Orb

    uid = Orb.comm_unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, 128, 0, device=device)
    ctx.comm_init(uid, rank, world)
    if peers:
        ctx.peer_import(all_gather_bytes(ctx.peer_export(), device=device), world)
