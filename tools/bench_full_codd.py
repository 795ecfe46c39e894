#!/usr/bin/env python
"""BASELINE.json configs[2]: full CODD (stereo + motion + fusion) 2-frame forward, 960x540 (padded 576x960), batch 4,
D=192, RAFT3D iters=16, random-init weights, through the reference-facing model(...) call.  Prints seconds per
2-frame sequence and the per-kernel table of one sequence (CUDA events per launch)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import codd_b200  # noqa: E402
from codd_b200 import ops  # noqa: E402
from codd_b200.synth import synth_pair  # noqa: E402


def main():
    B, H, W, D, iters = 4, 576, 960, 192, 16
    torch.manual_seed(0)
    model = codd_b200.build_estimator(codd_b200.codd_full_config(D, iters)).cuda()
    model.eval()
    left, right = synth_pair(B, H, W, D, seed=1234, kind="S")
    img = torch.stack([left, torch.roll(left, shifts=(1, 2), dims=(2, 3))], 1).cuda()
    r_img = torch.stack([right, torch.roll(right, shifts=(1, 2), dims=(2, 3))], 1).cuda()
    metas = [[dict(min_disp=1, max_disp=D, ori_shape=(540, 960), img_shape=(540, 960), intrinsics=[1050.0, 1050.0, 480.0, 270.0])]]

    def run():
        return model(return_loss=False, rescale=True, evaluate=False, img=[img], img_metas=metas, r_img=[r_img])[0]

    with torch.no_grad():
        for _ in range(2):
            out = run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            out = run()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        n0 = ops.LAUNCHES[0]
        with ops.profile() as prof:
            run()
        kernels = prof.summary()
    total = sum(k["ms"] for k in kernels)
    print(json.dumps({"config": "full CODD 2-frame forward, batch 4, 960x540 (576x960), D=192, iters=16",
                      "seconds_per_sequence": round(dt, 4), "frames_per_sec": round(2 * B / dt, 2),
                      "finite": bool(torch.isfinite(out).all()), "launches": ops.LAUNCHES[0] - n0,
                      "kernel_ms_total": round(total, 2),
                      "top_kernels": [{"kernel": k["kernel"], "ms": round(k["ms"], 3), "launches": k["launches"]}
                                      for k in kernels[:14]]}))


if __name__ == "__main__":
    main()
