"""One fused up+merge launch at the finest level (batch 16, 576x960, 16 channels) for ncu captures / timing:
   ncu --set full --import-source on --clock-control none -k regex:upmerge -s 2 -c 1 -o gpurun_out/upmerge python tools/upmerge_one.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from codd_b200 import ops
n, h, w = 16, 576, 960
coarse = ops.to_nhwc(torch.randn(n, 16, h // 2, w // 2, device="cuda"))
skip = ops.to_nhwc(torch.randn(n, 16, h, w, device="cuda"))
wu = ops.pack_deconv_weight(torch.randn(16, 16, 2, 2, device="cuda") / 8)
wm = ops.pack_conv_weight(torch.randn(16, 32, 1, 1, device="cuda") / 6)
bu, bm = torch.randn(16, device="cuda"), torch.randn(16, device="cuda")
for _ in range(3):
    ops.upmerge(coarse, skip, wu, bu, 16, wm, bm, 16)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.upmerge(coarse, skip, wu, bu, 16, wm, bm, 16)
e1.record()
torch.cuda.synchronize()
print("upmerge 16x576x960: %.1f us" % (e0.elapsed_time(e1) * 100))
