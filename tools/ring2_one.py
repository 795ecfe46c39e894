"""One fused two-convolution ring launch (16 -> 16 -> 16, batch 8, 576x960, residual) for ncu captures / timing:
   ncu --set full --import-source on --clock-control none -k regex:conv3x3x2 -s 2 -c 1 -o gpurun_out/ring2 python tools/ring2_one.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from codd_b200 import ops
n, h, w = 8, 576, 960
x = ops.to_nhwc(torch.randn(n, 16, h, w, device="cuda"))
wa = ops.pack_conv_weight_ring(torch.randn(16, 16, 3, 3, device="cuda") / 12)
wb = ops.pack_conv_weight_ring(torch.randn(16, 16, 3, 3, device="cuda") / 12)
ba, bb = torch.randn(16, device="cuda"), torch.randn(16, device="cuda")
for _ in range(3):
    ops.conv3x3x2_tc_ring(x, wa, ba, ops.ACT_LEAKY, wb, bb, ops.ACT_LEAKY, residual=x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.conv3x3x2_tc_ring(x, wa, ba, ops.ACT_LEAKY, wb, bb, ops.ACT_LEAKY, residual=x)
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10
print("ring2 8x16x576x960 (+res): %.1f us, %.0f GB/s algorithmic (in + res + out)" % (t * 1e3, 3 * 4 * n * h * w * 16 / t / 1e6))
