#!/usr/bin/env python
"""BASELINE.json configs[3]: full CODD (stereo + motion + fusion) sequence inference, num_frames=16, KITTI shape
1242x375 (padded 384x1280), D=192, RAFT3D iters=16, batch-sharded across the ranks of one box (reference:
inference.py:108-135 DistributedSampler + multi_gpu_inference, model/codd.py:290-398), through runner.run_model, i.e.
the reference-facing model(return_loss=False, ...) call.  One sequence batch per rank, no data-path collective.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 tools/bench_config3.py [--out f.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import codd_b200  # noqa: E402
from codd_b200 import ops  # noqa: E402
from codd_b200.runner import run_model  # noqa: E402
from codd_b200.sharding import broadcast_parameters, reduce_max_ms  # noqa: E402
from codd_b200.synth import synth_pair  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2, help="sequences per rank")
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    H, W, h, w, D, iters = 384, 1280, 375, 1242, 192, 16
    torch.manual_seed(0)
    model = codd_b200.build_estimator(codd_b200.codd_full_config(D, iters)).to(dev)
    model.eval()
    broadcast_parameters(model, src=0)
    left, right = synth_pair(a.batch, H, W, D, seed=77 + rank, kind="S")
    img = torch.stack([torch.roll(left, shifts=(t, 2 * t), dims=(2, 3)) for t in range(a.frames)], 1).to(dev)
    r_img = torch.stack([torch.roll(right, shifts=(t, 2 * t), dims=(2, 3)) for t in range(a.frames)], 1).to(dev)
    metas = [[dict(min_disp=1, max_disp=D, ori_shape=(h, w), img_shape=(h, w), intrinsics=[721.5, 721.5, w / 2.0, h / 2.0])]]
    out = run_model(model, [img], [r_img], metas)[0]          # warm-up (lazy weight packing, attribute set-up)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = ops.LAUNCHES[0]
    e0.record()
    for _ in range(a.reps):
        out = run_model(model, [img], [r_img], metas)[0]
    e1.record()
    torch.cuda.synchronize()
    ms = reduce_max_ms(e0.elapsed_time(e1), dev) / a.reps
    if rank == 0:
        line = {"config": "full CODD sequence inference, 16 frames, KITTI 1242x375 (384x1280), D=192, iters=16 (BASELINE.json configs[3])",
                "n_gpus": world, "sequences_per_gpu": a.batch, "frames": a.frames,
                "seconds_per_sequence_batch": round(ms * 1e-3, 4),
                "frames_per_sec": round(world * a.batch * a.frames / (ms * 1e-3), 2),
                "launches_per_sequence_batch": (ops.LAUNCHES[0] - n0) // a.reps,
                "output": list(out.shape), "finite": bool(torch.isfinite(out).all())}
        print(json.dumps(line))
        if a.out:
            json.dump(line, open(a.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
