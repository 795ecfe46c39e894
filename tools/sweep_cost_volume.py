#!/usr/bin/env python
"""BASELINE.json configs[4]: K1 sweep {540p, 720p, 1080p} x D in {64,128,192,256}, 8 samples per
GPU; achieved algorithmic GB/s (SURVEY.md §8d byte counts) of the materialising and the fused
arg-min variant, whole 5-level pyramid per measurement (ONE fused launch, codd_cost_volume_pyramid), CUDA events, L2 flushed between reps.

    python tools/sweep_cost_volume.py [--out profiles/cost_volume_sweep_rNN.json]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep_cost_volume.py ...

Under torchrun the batch of 64 of BASELINE.json configs[4] is sharded 8 samples per rank (no collective on the data path);
a row then reports the aggregate over ranks: bytes of all ranks / max-over-ranks time.  The fused variant is also given
against its own roof, the fp32 add pipe (2 * 16 * D lane-adds per (tile, disparity) pair with 4j - d >= 0).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from codd_b200 import ops  # noqa: E402

SIZES = {"540p": (576, 960), "720p": (768, 1280), "1080p": (1088, 1920)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch.distributed as dist
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    sm_mhz = 1965.0
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for name, (H, W) in SIZES.items():
        for D in (64, 128, 192, 256):
            levels = []
            for k in range(5):
                h, w = (H >> (4 - k)) // 4, (W >> (4 - k)) // 4
                tl = torch.randn(a.batch, 16, h, w, device=dev)          # planar, as K2 writes them
                tr = torch.randn(a.batch, 16, h, 4 * w, device=dev)
                levels.append((tl, tr, D // (16 >> k), h, w))
            for want_cv in (True, False):
                nbytes = sum(ops.cost_volume_bytes(a.batch, h, w, d, want_cv, True) for _, _, d, h, w in levels)
                ms = []
                for r in range(a.reps + 2):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    ops.cost_volume_pyramid([(tl, tr) for tl, tr, _, _, _ in levels], [d for _, _, d, _, _ in levels],
                                            want_cv=want_cv, want_argmin=True)
                    e1.record()
                    torch.cuda.synchronize()
                    if r >= 2:
                        ms.append(e0.elapsed_time(e1))
                t = sorted(ms)[len(ms) // 2]
                if world > 1:
                    tt = torch.tensor([t], dtype=torch.float64, device=dev)
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    t = tt.item()
                gbs = world * nbytes / (t * 1e-3) / 1e9
                row = dict(size=name, padded=[H, W], D=D, batch=a.batch * world, n_gpus=world,
                           variant="materialise+argmin" if want_cv else "fused argmin",
                           algorithmic_mb=round(world * nbytes / 1e6, 2), ms=round(t, 4), gbs=round(gbs, 1),
                           gbs_per_gpu=round(gbs / world, 1), frac_of_measured_peak=round(gbs / world / peak, 4))
                if not want_cv:
                    lane_adds = 0
                    for _, _, d, h, w in levels:
                        lane_adds += 2 * 16 * a.batch * h * sum(min(d, 4 * j + 1) for j in range(w))
                    row["fp32_roof_frac"] = round(lane_adds / (t * 1e-3) / (148 * 128 * sm_mhz * 1e6), 4)
                rows.append(row)
                if rank == 0:
                    print(row)
    if a.out and rank == 0:
        json.dump(dict(peak_gbs=peak, n_gpus=world, batch_per_gpu=a.batch, fp32_peak="148 SMs x 128 lanes x 1965 MHz",
                       rows=rows), open(a.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
