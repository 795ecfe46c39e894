"""Synthetic stereo inputs for benchmarks and smoke runs (Sets U / S / G).

Product-side copy of the generator (the oracle keeps its own so that neither side imports the
other); both produce identical tensors for identical arguments."""
import torch
import torch.nn.functional as F


def synth_pair(n, h, w, max_disp, seed=1234, kind="S"):
    """Synthetic stereo pair [N,3,h,w] x2 (h, w already multiples of 64).

    U: uniform [0,1), right == left (benchmark_speed.py:40-42; degenerate, timing only).
    S: band-limited texture in ImageNet-normalised range, right = left shifted by a smooth
       disparity field in [0, 0.8*max_disp)  (parity set).
    G: iid normal (kernel-level tests)."""
    g = torch.Generator().manual_seed(seed)
    if kind == "U":
        left = torch.rand(n, 3, h, w, generator=g)
        return left, left.clone()
    if kind == "G":
        return torch.randn(n, 3, h, w, generator=g), torch.randn(n, 3, h, w, generator=g)
    tex = torch.rand(n, 3, h, w + max_disp, generator=g)
    tex = F.avg_pool2d(F.pad(tex, (2, 2, 2, 2), mode="reflect"), 5, stride=1)
    tex = (tex - tex.mean()) / tex.std() * 1.1 + 0.2
    yy = torch.linspace(0, 1, h).view(1, h, 1)
    xx = torch.linspace(0, 1, w).view(1, 1, w)
    ph = torch.rand(n, 1, 1, generator=g) * 6.28
    d = 0.8 * max_disp * (0.5 + 0.25 * torch.sin(3.1 * xx + ph) * torch.cos(2.3 * yy) + 0.2 * yy)
    d = d.clamp(0, 0.8 * max_disp)
    base = torch.arange(w, dtype=torch.float32).view(1, 1, w) + max_disp
    left = _sample_x(tex, base.expand(n, h, w))
    right = _sample_x(tex, base + d)      # right[x] = left[x + d]  <=>  left[x] = right[x - d]
    return left, right


def _sample_x(tex, xs):
    n, c, h, wt = tex.shape
    x0 = xs.floor().clamp(0, wt - 2)
    f = (xs - x0).unsqueeze(1)
    i0 = x0.long().unsqueeze(1).expand(n, c, h, xs.shape[2])
    return torch.gather(tex, 3, i0) * (1 - f) + torch.gather(tex, 3, i0 + 1) * f
