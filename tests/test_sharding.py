"""Multi-GPU plumbing on CPU: the batch-sharding / weight-broadcast / max-over-ranks timing logic of
bench.py, exercised with a world of 2 gloo ranks (no CUDA, no kernels)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import codd_b200
    from codd_b200.sharding import broadcast_parameters, reduce_max_ms, shard_batch
    torch.manual_seed(100 + rank)                      # ranks start with DIFFERENT weights
    model = codd_b200.MODELS.build(codd_b200.hitnet_config(64))
    broadcast_parameters(model, src=0)                 # one flat broadcast (the reference's DDP wrap)
    flat = torch.cat([p.detach().flatten() for p in model.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    # buffers travel too (BatchNorm running statistics of the motion network's HRNet; int64 num_batches_tracked)
    bn = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 1), torch.nn.BatchNorm2d(4))
    bn[1].running_mean.fill_(float(rank + 1))
    bn[1].num_batches_tracked.fill_(10 * (rank + 1))
    broadcast_parameters(bn, src=0)
    same = same and bool((bn[1].running_mean == 1.0).all()) and int(bn[1].num_batches_tracked) == 10
    idx = shard_batch(10, rank, world)                 # rank r takes items r::world (DistributedSampler order)
    ms = reduce_max_ms(float(rank + 1) * 3.0)          # timing is the max over ranks
    q.put((rank, same, idx, ms))
    dist.destroy_process_group()


def test_world_size_two_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "weights differ after broadcast"
    assert res[0][2] == [0, 2, 4, 6, 8] and res[1][2] == [1, 3, 5, 7, 9]
    assert res[0][3] == res[1][3] == 6.0


def _metrics_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import numpy as np
    from codd_b200.metrics import ROW, gather_rows, summarise
    g = np.random.default_rng(7)
    rows_all = g.uniform(1, 50, (7, ROW))
    rows_all[3, 0] = 0.0                                  # a frame without valid pixels does not update the EPE meter
    mine = torch.from_numpy(rows_all[rank::world].copy())  # ranks hold different numbers of frames (4 and 3)
    got = summarise(gather_rows(mine).numpy())
    ref = summarise(np.concatenate([rows_all[r::world] for r in range(world)], 0))
    q.put((rank, got == ref, got["epe"]))
    dist.destroy_process_group()


def test_metric_rows_gathered_over_ranks():
    """N2/N3 on several GPUs: every rank evaluates its own sequences, the accumulator rows are all-gathered once and
    every rank reports the statistics of the whole dataset (apis/inference.py collect_results)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_metrics_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res) and res[0][2] == res[1][2] > 0
