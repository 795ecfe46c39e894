"""One rolling-ring conv launch for ncu captures: RC=16 -> 8x16x576x960, RC=32 -> 8x32x288x480.
   ncu --set full --import-source on --clock-control none -k regex:conv3x3_tc_ring -s 2 -c 1 -o gpurun_out/ring16 python tools/ring_one.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from codd_b200 import ops
n, c, h, w = 8, int(os.environ.get("RC", "16")), 576, 960
if c == 32: h, w = 288, 480
x = ops.to_nhwc(torch.randn(n, c, h, w, device="cuda"))
wt = torch.randn(c, c, 3, 3, device="cuda") / (c * 9) ** 0.5
b = torch.randn(c, device="cuda")
wr = ops.pack_conv_weight_ring(wt)
for _ in range(3):
    ops.conv3x3_tc_ring(x, wr, b, c, ops.ACT_LEAKY)
torch.cuda.synchronize()
