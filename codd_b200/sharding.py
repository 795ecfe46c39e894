"""Multi-GPU plumbing: one process per GPU, batch / sequence sharding, weights-only collectives.

Frames of one sequence are strictly sequential through ``state`` (model/codd.py:322-378) but batch
items / sequences are independent, so the path shards on the batch axis with NO data-path
collective — the reference does the same with a DistributedSampler (inference.py:108-115) and a DDP
wrap whose only run-time traffic in eval is the initial parameter broadcast (inference.py:130-134).
"""
import torch
import torch.distributed as dist


def shard_batch(n_items, rank, world):
    """Indices of the items rank ``rank`` processes: r, r+world, ... (DistributedSampler order)."""
    return list(range(rank, n_items, world))


def broadcast_parameters(model, src=0):
    """One flat broadcast of every parameter (NCCL on GPUs, gloo in the CPU tests)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    params = [p for p in model.parameters()]
    flat = torch.cat([p.data.flatten() for p in params])
    dist.broadcast(flat, src)
    off = 0
    for p in params:
        p.data.copy_(flat[off:off + p.numel()].view_as(p))
        off += p.numel()


def reduce_max_ms(ms, device=None):
    """Device-timed milliseconds -> max over ranks (the number a multi-GPU bench line reports)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()
