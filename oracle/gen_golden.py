"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (CPU, this container).

Run from the repo root where /root/reference is mounted:
    python -m oracle.gen_golden
The reference ships no golden vectors; these fixtures are the travelling pin:
they let the GPU box (where /root/reference does not exist) check both the oracle restatement
and the CUDA path against numbers the reference itself produced.
"""
import os
import sys
import warnings

import numpy as np
import torch

from . import hitnet_oracle as O
from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def hitnet_fixture(name, n, h, w, max_disp, wseed, dseed, kind, full=True):
    warnings.filterwarnings("ignore")
    model = ref_loader.build_hitnet(max_disp)
    sd = O.random_hitnet_params(wseed)
    model.load_state_dict(sd, strict=True)
    left, right = O.synth_pair(n, h, w, max_disp, seed=dseed, kind=kind)
    ns = ref_loader.load()
    with torch.no_grad():
        out = model.stereo_matching(left, right)
        fl = model.backbone(left)
        fr = model.backbone(right)
        tiles = model.tile_init.tile_features(fl, fr)
        cvs, hyps = model.tile_init.tile_hypothesis_pyramid(tiles, fl)
        tu = model.tile_update
        t16 = tu.tile_update0(fl[0], fr[0], hyps[0])
        t8 = tu.tile_update1(fl[1], fr[1], hyps[1], t16[0])
        # the raw local cost volume of the finest level (TileWarping, propagation.py:61-86)
        lcv = tu.tile_update4.tile_warping(hyps[4][:, :3], fl[4], fr[4])
        up = ns.propagation.upsample(t16[0], 2)
    fx = dict(
        meta=np.array([n, h, w, max_disp, wseed, dseed], dtype=np.int64),
        left=left.numpy(), right=right.numpy(),
        pred_disp=out["pred_disp"].numpy(), left_feat=out["left_feat"].numpy(),
        right_feat=out["right_feat"].numpy(),
        refined16=t16[0].numpy(), refined8=t8[0].numpy(), local_cv_l4=lcv.numpy(), upsample16=up.numpy(),
    )
    for k in range(5):
        fx[f"hyp{k}"] = hyps[k].numpy()
        if full or k < 3:   # the big fixture keeps only the coarse levels' intermediates
            fx[f"tile_l{k}"] = tiles[k][0].numpy()
            fx[f"tile_r{k}"] = tiles[k][1].numpy()
            fx[f"cv{k}"] = cvs[k].numpy()
        if full:
            fx[f"fea_l{k}"] = fl[k].numpy()
    if not full:
        del fx["local_cv_l4"]
    # weights are re-derivable from wseed; a digest of the raw bytes guards against RNG drift
    fx["weights_sha1"] = np.frombuffer(O.params_digest(sd), dtype=np.uint8)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **fx)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def main():
    if not ref_loader.available():
        sys.exit("reference not mounted; fixtures can only be generated where /root/reference exists")
    hitnet_fixture("hitnet_s_128x128_d32.npz", 1, 128, 128, 32, wseed=7, dseed=11, kind="S")
    hitnet_fixture("hitnet_s_128x192_d64.npz", 1, 128, 192, 64, wseed=3, dseed=5, kind="S", full=False)


if __name__ == "__main__":
    main()
