"""In-graph cost of the coarse-level 3x3 convolutions (halo-tile tcgen05 kernel): 17 launches as in one stereo step
(TileUpdate levels 0-2 and the backbone's coarsest block), replayed from a CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from codd_b200 import ops
dev = "cuda"
shapes = [(8, 36, 60)] * 5 + [(8, 18, 30)] * 5 + [(8, 9, 15)] * 4 + [(16, 18, 30)] * 3
wt = torch.randn(32, 32, 3, 3, device=dev) / 17
b = torch.randn(32, device=dev)
ws = ops.pack_conv_weight_tc(wt)
xs = [ops.to_nhwc(torch.randn(n, 32, h, w, device=dev)) for (n, h, w) in shapes]
def run():
    for x in xs:
        ops.conv3x3_tc(x, ws, b, 32, ops.ACT_LEAKY)
for _ in range(3):
    run()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    run()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    g.replay()
e0.record()
for _ in range(20):
    g.replay()
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 20
print("17 coarse-level 3x3 convs in a graph: %.1f us total, %.1f us per launch" % (t * 1e3, t * 1e3 / len(shapes)))
