#!/usr/bin/env python
"""K4 timing probe at the bench's level-0 geometry (N=8, 576x960 features, C=16, 144x240 tiles), for smooth and
noisy hypothesis fields and several staging thresholds (CODD_K4_MAXWIN is read once per process -> one run each)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from codd_b200 import ops  # noqa: E402


def main():
    n, c, h, w = 8, 16, 144, 240
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    fl = ops.to_nhwc(torch.randn(n, c, 4 * h, 4 * w, device=dev, generator=g))
    fr = torch.randn(n, c, 4 * h, 4 * w, device=dev, generator=g)
    if os.environ.get("CODD_K4_NHWC", "1") != "0":
        fr = ops.to_nhwc(fr)      # gathered in place (codd_tile_warp_cost_nhwc)
    dec_w = torch.randn(16, 64, device=dev, generator=g) / 8
    dec_b = torch.zeros(16, device=dev)
    yy, xx = torch.meshgrid(torch.arange(h, device=dev).float(), torch.arange(w, device=dev).float(), indexing="ij")
    kinds = os.environ.get("K4_KINDS", "init,smooth,noisy").split(",")
    for kind in kinds:
        cur = torch.zeros(n, 16, h, w, device=dev)
        prev = torch.zeros(n, 16, h // 2, w // 2, device=dev)
        if kind in ("smooth", "init"):
            cur[:, 0] = 60 + 40 * torch.sin(xx / 37) * torch.cos(yy / 23)
            prev[:, 0] = (cur[:, 0, ::2, ::2] + 0.7) / 2
            prev[:, 1:3] = torch.randn(n, 2, h // 2, w // 2, device=dev, generator=g) * 0.05
        else:
            cur[:, 0] = torch.rand(n, h, w, device=dev, generator=g) * 190
            prev[:, 0] = torch.rand(n, h // 2, w // 2, device=dev, generator=g) * 95
        cur[:, 1:3] = torch.randn(n, 2, h, w, device=dev, generator=g) * 0.2
        if kind == "init":       # what the network feeds: arg-min initialisations (integer disparity, zero slants)
            cur[:, 0] = cur[:, 0].round()
            cur[:, 1:3] = 0.0
        cur_n, prev_n = ops.to_nhwc(cur), ops.to_nhwc(prev)
        for _ in range(3):
            ops.tile_warp_cost(fl, fr, cur_n, prev_n, dec_w, dec_b)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.tile_warp_cost(fl, fr, cur_n, prev_n, dec_w, dec_b)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gb = ops.tile_warp_bytes(n, c, h, w, True) / 1e9
        print(f"right={'nhwc' if not fr.is_contiguous() else 'planar'} maxwin={os.environ.get('CODD_K4_MAXWIN', 'default')} {kind}: {ms:.3f} ms  {gb / (ms * 1e-3):.0f} GB/s")


if __name__ == "__main__":
    main()
