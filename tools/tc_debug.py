#!/usr/bin/env python
"""Per-role wait-time breakdown of the tcgen05 conv (diagnostic counters)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from codd_b200 import lib, ops
from codd_b200.lib import ACT_LEAKY

for (cin, cout, n, h, w) in [(16, 16, 16, 576, 960), (32, 32, 8, 288, 480)]:
    x = ops.to_nhwc(torch.randn(n, cin, h, w, device="cuda"))
    wt = torch.randn(cout, cin, 3, 3, device="cuda") / (cin * 9) ** 0.5
    b = torch.randn(cout, device="cuda")
    ws = ops.pack_conv_weight_tc(wt)
    dbg = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
    for _ in range(3):
        ops.conv3x3_tc(x, ws, b, cout, ACT_LEAKY)
    __import__('ctypes').CDLL(lib.LIB_PATH).codd_conv3x3_tc_debug(__import__('ctypes').c_void_p(dbg.data_ptr()))   # needs a `make DIAG=1` build
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for fl in (4, 8, 12):
        e0.record(); ops.conv3x3_tc(x, ws, b, cout, ACT_LEAKY, flags=fl); e1.record(); torch.cuda.synchronize()
        print(f"   diag flags {fl} (4=no stores, 8=no split, 12=neither): {e0.elapsed_time(e1)*1e3:.0f} us")
    e0.record(); ops.conv3x3_tc(x, ws, b, cout, ACT_LEAKY); e1.record(); torch.cuda.synchronize()
    __import__('ctypes').CDLL(lib.LIB_PATH).codd_conv3x3_tc_debug(None)
    d = dbg.view(148, 8).double().mean(0).tolist()
    names = ["epi tmem-ld cycles", "mma wait-full", "mma wait-acc-empty", "mma wait-lo", "epi wait-acc-full",
             "split wait-p12", "split work", "epi math+store cycles"]
    tiles = n * ((h + 1) // 2) * ((w + 127) // 128) / 148
    print(f"cin={cin} cout={cout}: {e0.elapsed_time(e1)*1e3:.0f} us, {tiles:.0f} tiles/SM; mean cycles per tile:")
    for nm, v in zip(names, d):
        print(f"   {nm:22s} {v / tiles:9.0f}")
