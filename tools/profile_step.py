#!/usr/bin/env python
"""Runs a few eager (no CUDA graph) HITNetMF steps at the bench configuration, for ncu:
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c <n> \
        -o gpurun_out/prof python tools/profile_step.py --steps 1
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import codd_b200  # noqa: E402
from codd_b200.synth import synth_pair  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--hw", type=int, nargs=2, default=[576, 960])
    ap.add_argument("--max-disp", type=int, default=192)
    ap.add_argument("--cv", action="store_true", help="also run the materialising cost-volume variant per level")
    ap.add_argument("--tags", default=None, help="write the launch-order list of ops tags of the LAST step (JSON): "
                                                 "joined with an ncu launch list by tools/traffic_table.py")
    a = ap.parse_args()
    torch.manual_seed(0)
    m = codd_b200.MODELS.build(codd_b200.hitnet_config(a.max_disp)).cuda().eval()
    left, right = synth_pair(a.batch, a.hw[0], a.hw[1], a.max_disp, seed=1234, kind="S")
    left, right = left.cuda(), right.cuda()
    from codd_b200 import ops
    with torch.no_grad():
        for i in range(a.steps):
            if a.tags and i == a.steps - 1:
                with ops.profile() as prof:
                    out = m.stereo_matching(left, right)
                import json
                json.dump([[t, b] for t, b, _, _ in prof.records], open(a.tags, "w"))
            else:
                out = m.stereo_matching(left, right)
        if a.cv:
            m.tile_init.materialize_cv = True
            fl, fr = m.backbone.forward_pair(left, right)
            m.tile_init(fl, fr)
    torch.cuda.synchronize()
    print("done", tuple(out["pred_disp"].shape))


if __name__ == "__main__":
    main()
