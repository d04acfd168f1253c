"""One-process-per-GPU plumbing (torch.distributed is only the rendezvous; the data-path
collectives are NCCL calls made by the CUDA library on its own stream).

Row sharding follows SURVEY.md section 8e: contiguous row blocks, rank r owns
[r * rows_per_rank, min(n, (r + 1) * rows_per_rank)), rows_per_rank rounded up to the 1024-row
kmeans++ summation block so a shard boundary never splits a block.
"""
import os

SHARD_ALIGN = 1024


def shard_range(n, world, rank, align=SHARD_ALIGN):
    """Half-open global row range of `rank`.  Deterministic, covers [0, n) exactly once."""
    per = -(-n // world)
    per = -(-per // align) * align
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi


def env_world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def broadcast_bytes(payload, src=0, length=128):
    """Ship a small byte string (the ncclUniqueId) from `src` to every rank of the default group."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(length, dtype=torch.uint8)
    if dist.get_rank() == src:
        t = torch.tensor(list(payload), dtype=torch.uint8)
    dev = None
    if dist.get_backend() == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device()); t = t.to(dev)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().tolist())


def max_over_ranks(value):
    """max of a python float over the default group (device timing: max over ranks)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.to(torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def join_comm(ctx):
    """Create the library's NCCL communicator across the ranks of the default process group."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    uid = ctx.comm_unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, 0, 128)
    ctx.comm_init_rank(world, rank, uid)
    return world, rank
