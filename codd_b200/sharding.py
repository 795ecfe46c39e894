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
    """Module state of rank `src` to every rank: parameters AND buffers (the reference's DDP wrap syncs both,
    inference.py:130-134 — HRNet's BatchNorm running statistics are buffers, and they are folded into the packed
    convolution weights), one flat broadcast per (dtype, device) group (NCCL on GPUs, gloo in the CPU tests)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    groups = {}
    for t in list(model.parameters()) + list(model.buffers()):
        groups.setdefault((t.dtype, t.device), []).append(t)
    for (dtype, _), tensors in groups.items():
        if dtype == torch.bool:        # no broadcast kernel for bool on every backend
            flat = torch.cat([t.data.flatten().to(torch.uint8) for t in tensors])
        else:
            flat = torch.cat([t.data.flatten() for t in tensors])
        dist.broadcast(flat, src)
        off = 0
        for t in tensors:
            t.data.copy_(flat[off:off + t.numel()].view_as(t).to(t.dtype))
            off += t.numel()


def reduce_max_ms(ms, device=None):
    """Device-timed milliseconds -> max over ranks (the number a multi-GPU bench line reports)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()
