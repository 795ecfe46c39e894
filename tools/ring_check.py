import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from codd_b200 import ops
from codd_b200.lib import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_RELU_CH0
torch.manual_seed(0)
def act_ref(v, act):
    if act == ACT_LEAKY: return F.leaky_relu(v, 0.2)
    if act == ACT_RELU: return F.relu(v)
    if act == ACT_RELU_CH0:
        v = v.clone(); v[:, 0] = F.relu(v[:, 0]); return v
    return v
for (n, cin, cout, h, w) in [(2,16,16,128,192),(4,16,16,128,192),(2,32,32,64,96),(2,24,24,64,96),(2,16,1,128,192),(2,32,16,64,96),(2,32,32,32,48),(1,16,16,128,192),(2,16,16,128,64)]:
    for act in (ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_RELU_CH0):
        for resmode in ("none", "full", "bcast"):
            x = torch.randn(n, cin, h, w) * 3
            wt = torch.randn(cout, cin, 3, 3) / (cin * 9) ** 0.5
            b = torch.randn(cout)
            res = None if resmode == "none" else torch.randn(n, cout if resmode == "full" else 1, h, w)
            ref = F.conv2d(x, wt, b, padding=1)
            if res is not None: ref = ref + res
            ref = act_ref(ref, act)
            out = ops.conv3x3_tc_ring(ops.to_nhwc(x.cuda()), ops.pack_conv_weight_ring(wt.cuda()), b.cuda(), cout, act,
                                      residual=None if res is None else ops.to_nhwc(res.cuda()), res_bcast=resmode == "bcast")
            err = (ops.to_nchw(out).cpu() - ref).abs().max().item()
            if err > 1e-4: print("BAD", (n, cin, cout, h, w), act, resmode, err)
print("done")
