#!/usr/bin/env python
"""Timing probe: 3x3 conv kernels (tcgen05 halo-tile vs rolling ring) at the bench's dominant shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from codd_b200 import ops  # noqa: E402


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda")
    for (n, c, h, w) in [(16, 16, 576, 960), (8, 16, 576, 960), (8, 32, 288, 480), (8, 32, 144, 240), (8, 32, 36, 60)]:
        x = ops.to_nhwc(torch.randn(n, c, h, w, device=dev))
        wt = torch.randn(c, c, 3, 3, device=dev) / (c * 9) ** 0.5
        b = torch.randn(c, device=dev)
        ws, wr = ops.pack_conv_weight_tc(wt), ops.pack_conv_weight_ring(wt)
        t_old = timeit(lambda: ops.conv3x3_tc(x, ws, b, c, ops.ACT_LEAKY))
        t_new = timeit(lambda: ops.conv3x3_tc_ring(x, wr, b, c, ops.ACT_LEAKY))
        gb = 4 * n * h * w * 2 * c / 1e9
        import ctypes
        from codd_b200 import lib
        hook = getattr(ctypes.CDLL(lib.LIB_PATH), "codd_conv3x3_tc_ring_debug", None)   # only in `make DIAG=1` builds
        if hook is not None:
            dbg = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
            hook(ctypes.c_void_p(dbg.data_ptr()))
            ops.conv3x3_tc_ring(x, wr, b, c, ops.ACT_LEAKY)
            torch.cuda.synchronize()
            hook(None)
            d = dbg.view(148, 8).float().mean(0).tolist()
            print("   ring role waits (mean clk/CTA): producer-empty %.0f | A: full %.0f | B: lo %.0f | A: slot %.0f total %.0f | B: iss %.0f | "
                  "epi accf %.0f | split p12 %.0f" % tuple(d))
        print(f"N={n} C={c} {h}x{w}: halo-tile {t_old * 1e3:.1f} us ({gb / t_old * 1e3:.0f} GB/s)   ring {t_new * 1e3:.1f} us "
              f"({gb / t_new * 1e3:.0f} GB/s)")


if __name__ == "__main__":
    main()
