"""Kernel-level parity: every C-ABI entry point against the CPU oracle on the same seeded
inputs.  Integer / index / discrete outputs and the whole cost-volume + warp arithmetic are
compared BIT-EXACT; the dense convolutions (fp32 FMA, different summation order than torch's
CPU kernels) within rtol 2e-5 / atol 2e-5, tolerance stated per test."""
import numpy as np
import os

import pytest
import torch
import torch.nn.functional as F

from conftest import golden_params
from oracle import hitnet_oracle as O

pytestmark = pytest.mark.gpu

CONV_RTOL, CONV_ATOL = 2e-5, 2e-5


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from codd_b200 import ops as _ops
    return _ops


def dev(x):
    return x.cuda()


def nhwc(ops, x):
    return ops.to_nhwc(x.cuda().float())


def back(ops, t):
    return ops.to_nchw(t).cpu()


def gen(seed):
    return torch.Generator().manual_seed(seed)


# ------------------------------------------------------------------------------------------
def test_layout_roundtrip(ops):
    x = torch.randn(3, 24, 7, 13, generator=gen(0))
    t = nhwc(ops, x)
    assert t.shape == x.shape and ops.ld_of(t) == 24
    assert torch.equal(back(ops, t), x)
    assert torch.equal(back(ops, t[:, 8:16]), x[:, 8:16])
    assert torch.equal(back(ops, t[1:]), x[1:])


CONV_CASES = [
    # k, stride, pad, dil, cin, cout, h, w
    (1, 1, 0, 1, 16, 16, 9, 15), (1, 1, 0, 1, 48, 24, 18, 30), (1, 1, 0, 1, 64, 32, 5, 70), (1, 1, 0, 1, 32, 13, 8, 8),
    (3, 1, 1, 1, 16, 16, 36, 60), (3, 1, 1, 1, 24, 24, 17, 33), (3, 1, 1, 1, 32, 32, 40, 40), (3, 1, 1, 1, 32, 34, 9, 15),
    (3, 1, 1, 1, 16, 3, 20, 50), (3, 1, 1, 1, 16, 1, 20, 50), (3, 1, 1, 1, 32, 16, 33, 31), (3, 1, 1, 1, 8, 8, 6, 6),
    (3, 1, 3, 3, 32, 32, 36, 60), (4, 2, 1, 1, 16, 16, 32, 64), (4, 2, 1, 1, 16, 24, 36, 60), (4, 2, 1, 1, 24, 32, 18, 30),
    (4, 4, 0, 1, 16, 16, 36, 60), (4, 4, 0, 1, 32, 16, 8, 140), (7, 1, 3, 1, 2, 32, 20, 30), (7, 1, 3, 1, 64, 30, 12, 34),
    (1, 1, 0, 1, 31, 64, 10, 10), (3, 1, 1, 1, 17, 5, 7, 9), (3, 1, 1, 1, 32, 2, 9, 15), (3, 1, 1, 1, 32, 2, 37, 70),
]


@pytest.mark.parametrize("k,s,p,d,cin,cout,h,w", CONV_CASES)
def test_conv2d_vs_torch_cpu(ops, k, s, p, d, cin, cout, h, w):
    from codd_b200.lib import ACT_LEAKY
    g = gen(k * 1000 + cin * 10 + cout)
    x = torch.randn(2, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.leaky_relu(F.conv2d(x, wt, b, stride=s, padding=p, dilation=d), 0.2)
    out = ops.conv2d(nhwc(ops, x), ops.pack_conv_weight(wt).cuda(), b.cuda(), cout, k, s, p, d, ACT_LEAKY)
    assert out.shape == ref.shape
    torch.testing.assert_close(back(ops, out), ref, rtol=CONV_RTOL, atol=CONV_ATOL)


def test_conv3x3_two_filter_head_into_slice(ops):
    """The two confidence filters of TileUpdate.lastconv (32 -> 2 of a 34-channel layer) through the two-output head kernel,
    written into channels [32, 34) of a 36-wide buffer, weights a column slice semantics of the packed layer."""
    from codd_b200.lib import ACT_NONE
    g = gen(77)
    x = torch.randn(2, 32, 21, 45, generator=g)
    wt = torch.randn(34, 32, 3, 3, generator=g) / 17
    b = torch.randn(34, generator=g)
    ref = F.conv2d(x, wt, b, padding=1)[:, 32:34]
    out = ops.empty_nhwc(2, 34, 21, 45, "cuda", 36)
    out.zero_()
    ops.conv2d(nhwc(ops, x), ops.pack_conv_weight(wt[32:34]).cuda(), b[32:34].cuda().contiguous(), 2, 3, 1, 1, 1, ACT_NONE,
               out=out[:, 32:34])
    torch.testing.assert_close(back(ops, out[:, 32:34]), ref, rtol=CONV_RTOL, atol=CONV_ATOL)
    assert float(out[:, :32].abs().max()) == 0.0


def test_conv2d_dual_input_residual_slices(ops):
    from codd_b200.lib import ACT_NONE, ACT_RELU, ACT_RELU_CH0
    g = gen(5)
    a = torch.randn(2, 24, 19, 37, generator=g)
    b2 = torch.randn(2, 16, 19, 37, generator=g)
    wt = torch.randn(32, 40, 1, 1, generator=g) / 6
    bias = torch.randn(32, generator=g)
    ref = F.conv2d(torch.cat([a, b2], 1), wt, bias)
    # a lives as a channel slice [8:32] of a 48-wide buffer
    wide = ops.empty_nhwc(2, 48, 19, 37, "cuda")
    wide[:, 8:32].copy_(a.cuda())
    out = ops.conv2d(wide[:, 8:32], ops.pack_conv_weight(wt).cuda(), bias.cuda(), 32, 1, act=ACT_NONE,
                     x2=nhwc(ops, b2))
    torch.testing.assert_close(back(ops, out), ref, rtol=CONV_RTOL, atol=CONV_ATOL)
    # residual (full) + relu on channel 0 only, 3x3
    x = torch.randn(2, 32, 19, 37, generator=g)
    w3 = torch.randn(16, 32, 3, 3, generator=g) / 17
    b3 = torch.randn(16, generator=g)
    res = torch.randn(2, 16, 19, 37, generator=g)
    r = F.conv2d(x, w3, b3, padding=1) + res
    r = torch.cat([F.relu(r[:, :1]), r[:, 1:]], 1)
    out = ops.conv2d(nhwc(ops, x), ops.pack_conv_weight(w3).cuda(), b3.cuda(), 16, 3, 1, 1, 1, ACT_RELU_CH0,
                     residual=nhwc(ops, res))
    torch.testing.assert_close(back(ops, out), r, rtol=CONV_RTOL, atol=CONV_ATOL)
    # broadcast single-channel residual + relu
    r = F.relu(F.conv2d(x, w3[:3], b3[:3], padding=1) + res[:, :1])
    out = ops.conv2d(nhwc(ops, x), ops.pack_conv_weight(w3[:3]).cuda(), b3[:3].cuda(), 3, 3, 1, 1, 1, ACT_RELU,
                     residual=nhwc(ops, res)[:, :1], res_bcast=True)
    torch.testing.assert_close(back(ops, out), r, rtol=CONV_RTOL, atol=CONV_ATOL)


def test_conv3x3_head1_residual_relu(ops):
    """Single-channel head kernel (propagation.py:325-333 in eval mode): ragged tile edges, 1-channel residual,
    relu, input as a channel slice of a wider buffer."""
    from codd_b200.lib import ACT_RELU
    g = gen(11)
    x = torch.randn(3, 16, 37, 71, generator=g)
    wt = torch.randn(1, 16, 3, 3, generator=g) / 12
    b = torch.randn(1, generator=g)
    res = torch.randn(3, 1, 37, 71, generator=g)
    ref = F.relu(F.conv2d(x, wt, b, padding=1) + res)
    wide = ops.empty_nhwc(3, 32, 37, 71, "cuda")
    wide[:, 16:32].copy_(x.cuda())
    out = ops.conv2d(wide[:, 16:32], ops.pack_conv_weight(wt).cuda(), b.cuda(), 1, 3, 1, 1, 1, ACT_RELU,
                     residual=nhwc(ops, res), res_bcast=True)
    torch.testing.assert_close(back(ops, out), ref, rtol=CONV_RTOL, atol=CONV_ATOL)


@pytest.mark.parametrize("cc,cs,cu,co,n,h,w", [(16, 16, 16, 16, 2, 20, 600), (24, 16, 16, 16, 1, 6, 1030), (24, 24, 24, 24, 2, 10, 70),
                                                (32, 24, 24, 24, 1, 4, 514), (16, 16, 16, 16, 1, 2, 2), (16, 16, 16, 16, 3, 6, 1000),
                                                (16, 16, 16, 16, 1, 4, 192)])
def test_upmerge_fused(ops, cc, cs, cu, co, n, h, w):
    """conv_up + conv_merge[0] in one kernel (backbone.py:17-32,75-88) against the two torch ops it replaces."""
    g = gen(cc * 10 + cs + h + w)
    coarse = torch.randn(n, cc, h // 2, w // 2, generator=g)
    skip = torch.randn(n, cs, h, w, generator=g)
    wu = torch.randn(cc, cu, 2, 2, generator=g) / cc ** 0.5
    bu = torch.randn(cu, generator=g)
    wm = torch.randn(co, cs + cu, 1, 1, generator=g) / (cs + cu) ** 0.5
    bm = torch.randn(co, generator=g)
    up = F.leaky_relu(F.conv_transpose2d(coarse, wu, bu, stride=2), 0.2)
    ref = F.leaky_relu(F.conv2d(torch.cat((skip, up), 1), wm, bm), 0.2)
    out = ops.upmerge(nhwc(ops, coarse), nhwc(ops, skip), ops.pack_deconv_weight(wu).cuda(), bu.cuda(), cu,
                      ops.pack_conv_weight(wm).cuda(), bm.cuda(), co)
    torch.testing.assert_close(back(ops, out), ref, rtol=CONV_RTOL, atol=CONV_ATOL)


@pytest.mark.parametrize("c,n,h,w", [(16, 2, 64, 256), (16, 1, 128, 132), (16, 1, 4, 4), (16, 2, 36, 520), (24, 2, 72, 120),
                                      (24, 1, 36, 264), (32, 2, 36, 60), (32, 1, 8, 520)])
def test_tile_features_tensor_core(ops, c, n, h, w):
    """K2 on tcgen05 (Cin = 16): left = 4x4 stride 4, right = stride (4,1) over the input zero-padded by 3 columns
    (initialization.py:119-124), LeakyReLU + 1x1 + LeakyReLU, planar output — against the oracle, conv tolerance."""
    g = gen(h * 7 + w + c)
    sd = {"t.0.weight": torch.randn(16, c, 4, 4, generator=g) / (16 * c) ** 0.5, "t.0.bias": torch.randn(16, generator=g),
          "t.2.weight": torch.randn(16, 16, 1, 1, generator=g) / 4.0, "t.2.bias": torch.randn(16, generator=g)}
    fl, fr = torch.randn(n, c, h, w, generator=g), torch.randn(n, c, h, w, generator=g)
    tl, tr = O.tile_features_level(sd, "t", fl, fr)
    ws = ops.pack_conv_weight_tc4(sd["t.0.weight"].cuda())
    w1 = sd["t.2.weight"].reshape(16, 16).cuda().contiguous()
    args = (ws, sd["t.0.bias"].cuda(), w1, sd["t.2.bias"].cuda())
    got_l = ops.tile_features_tc(nhwc(ops, fl), *args, right=False)
    got_r = ops.tile_features_tc(nhwc(ops, fr), *args, right=True)
    torch.cuda.synchronize()
    assert got_l.is_contiguous() and got_l.shape == tl.shape and got_r.shape == tr.shape
    print(f"tile features tc c={c} {n}x{h}x{w}: max abs err left {(got_l.cpu() - tl).abs().max().item():.3e} "
          f"right {(got_r.cpu() - tr).abs().max().item():.3e}")
    torch.testing.assert_close(got_l.cpu(), tl, rtol=CONV_RTOL, atol=CONV_ATOL)
    torch.testing.assert_close(got_r.cpu(), tr, rtol=CONV_RTOL, atol=CONV_ATOL)


def test_tile_conv_right_stride41(ops):
    """initialization.py:121-124: stride (4,1) over the input zero-padded 3 columns on the right."""
    from codd_b200.lib import ACT_LEAKY
    g = gen(8)
    x = torch.randn(2, 24, 16, 44, generator=g)
    wt = torch.randn(16, 24, 4, 4, generator=g) / 20
    b = torch.randn(16, generator=g)
    ref = F.leaky_relu(F.conv2d(F.pad(x, (0, 3, 0, 0)), wt, b, stride=(4, 1)), 0.2)
    out = ops.conv2d(nhwc(ops, x), ops.pack_conv_weight(wt).cuda(), b.cuda(), 16, 4, (4, 1), (0, 0), 1, ACT_LEAKY,
                     out_hw=(4, 44))
    assert out.shape == ref.shape
    torch.testing.assert_close(back(ops, out), ref, rtol=CONV_RTOL, atol=CONV_ATOL)


@pytest.mark.parametrize("cin,h,w", [(16, 8, 44), (24, 16, 132), (32, 4, 520), (16, 12, 2052)])
def test_tile_features_fused(ops, cin, h, w):
    """K2: conv4x4 (stride 4 / stride (4,1) + right zero pad 3) -> LReLU -> conv1x1 -> LReLU, planar out."""
    g = gen(cin + h + w)
    x = torch.randn(2, cin, h, w, generator=g)
    w0 = torch.randn(16, cin, 4, 4, generator=g) / (cin * 16) ** 0.5
    b0 = torch.randn(16, generator=g)
    w1 = torch.randn(16, 16, 1, 1, generator=g) / 4
    b1 = torch.randn(16, generator=g)
    sd = {"t.0.weight": w0, "t.0.bias": b0, "t.2.weight": w1, "t.2.bias": b1}
    ref_l, ref_r = O.tile_features_level(sd, "t", x, x)
    args = (ops.pack_conv_weight(w0).cuda(), b0.cuda(), w1.cuda().contiguous(), b1.cuda())
    out_l = ops.tile_features(nhwc(ops, x), *args, right=False)
    out_r = ops.tile_features(nhwc(ops, x), *args, right=True)
    assert out_l.is_contiguous() and out_l.shape == ref_l.shape and out_r.shape == ref_r.shape
    torch.testing.assert_close(out_l.cpu(), ref_l, rtol=CONV_RTOL, atol=CONV_ATOL)
    torch.testing.assert_close(out_r.cpu(), ref_r, rtol=CONV_RTOL, atol=CONV_ATOL)


@pytest.mark.parametrize("n,h,w", [(2, 21, 260), (1, 8, 64), (1, 40, 1000), (3, 9, 256), (1, 17, 132)])
def test_image_conv_tma_staged(ops, n, h, w):
    """conv1 of HITUNet (backbone.py:35-39) through the TMA-staged kernel (width a multiple of 4): ragged last CTA, image
    heights that are not a multiple of the 8-row tile, the zero padding delivered by TMA's out-of-bounds fill."""
    g = gen(n * 100 + h + w)
    l = torch.randn(n, 3, h, w, generator=g)
    r = torch.randn(n, 3, h, w, generator=g)
    wt = torch.randn(16, 3, 3, 3, generator=g) / 5
    b = torch.randn(16, generator=g)
    ref = F.leaky_relu(F.conv2d(torch.cat([l, r]), wt, b, padding=1), 0.2)
    out = ops.conv3x3_image(l.cuda(), r.cuda(), ops.pack_conv_weight(wt).cuda(), b.cuda(), 16)
    torch.testing.assert_close(back(ops, out), ref, rtol=CONV_RTOL, atol=CONV_ATOL)
    out1 = ops.conv3x3_image(l.cuda(), None, ops.pack_conv_weight(wt).cuda(), b.cuda(), 16)
    torch.testing.assert_close(back(ops, out1), ref[:n], rtol=CONV_RTOL, atol=CONV_ATOL)


def test_image_conv_and_deconv(ops):
    g = gen(9)
    l = torch.randn(2, 3, 21, 45, generator=g)
    r = torch.randn(2, 3, 21, 45, generator=g)
    wt = torch.randn(16, 3, 3, 3, generator=g) / 5
    b = torch.randn(16, generator=g)
    ref = F.leaky_relu(F.conv2d(torch.cat([l, r]), wt, b, padding=1), 0.2)
    out = ops.conv3x3_image(l.cuda(), r.cuda(), ops.pack_conv_weight(wt).cuda(), b.cuda(), 16)
    torch.testing.assert_close(back(ops, out), ref, rtol=CONV_RTOL, atol=CONV_ATOL)
    for cin, cout in ((32, 24), (24, 24), (24, 16), (16, 16)):
        x = torch.randn(2, cin, 9, 15, generator=g)
        wt = torch.randn(cin, cout, 2, 2, generator=g) / cin ** 0.5
        b = torch.randn(cout, generator=g)
        ref = F.leaky_relu(F.conv_transpose2d(x, wt, b, stride=2), 0.2)
        out = ops.deconv2x2(nhwc(ops, x), ops.pack_deconv_weight(wt).cuda(), b.cuda(), cout)
        torch.testing.assert_close(back(ops, out), ref, rtol=CONV_RTOL, atol=CONV_ATOL)


# ------------------------------------------------------------------------------------------
# K1: bit-exact
# ------------------------------------------------------------------------------------------
CV_CASES = [(1, 2, 2, 4), (1, 2, 5, 2), (2, 3, 9, 6), (1, 2, 40, 21), (1, 2, 3, 1), (1, 1, 600, 320), (2, 9, 15, 12), (1, 5, 33, 32), (2, 3, 40, 48), (1, 4, 70, 192), (1, 2, 240, 192),
            (1, 2, 300, 64), (1, 1, 500, 256), (1, 3, 7, 64)]


@pytest.mark.parametrize("n,h,w,d", CV_CASES)
def test_cost_volume_bit_exact(ops, n, h, w, d):
    g = gen(n + h * 7 + w * 13 + d)
    tl = torch.randn(n, 16, h, w, generator=g)
    tr = torch.randn(n, 16, h, 4 * w, generator=g)
    for j in range(w):                                   # plant true matches at varying disparities
        x = 4 * j - (j % 5) * 3
        if x >= 0:
            tr[:, :, :, x] = tl[:, :, :, j] + 0.01 * torch.randn(n, 16, h, generator=g)
    ref = O.cost_volume(tl, tr, d)
    rc, ri = ref.min(1)
    for want_cv, want_arg in ((True, True), (False, True), (True, False)):
        cv, mc, md = ops.cost_volume(nhwc(ops, tl), nhwc(ops, tr), d, want_cv=want_cv, want_argmin=want_arg)
        if want_cv:
            assert torch.equal(cv.cpu(), ref), "cost volume not bit-identical"
        if want_arg:
            assert torch.equal(mc.cpu()[:, 0], rc), "min cost not bit-identical"
            assert torch.equal(md.cpu()[:, 0], ri.float()), "arg-min indices differ"


def test_cost_volume_pyramid_matches_per_level(ops):
    """The fused 5-level launch is bit-identical to per-level launches (materialising and fused arg-min variants)."""
    n, D = 2, 64
    g = gen(77)
    tiles, disps = [], []
    for k in range(5):
        h, w = 3 << k, 5 << k
        tiles.append((torch.randn(n, 16, h, w, generator=g).cuda(), torch.randn(n, 16, h, 4 * w, generator=g).cuda()))
        disps.append(D // (16 >> k))
    for want_cv in (True, False):
        fused = ops.cost_volume_pyramid(tiles, disps, want_cv=want_cv)
        for (tl, tr), d, (cv, mc, md) in zip(tiles, disps, fused):
            cv1, mc1, md1 = ops.cost_volume(tl, tr, d, want_cv=want_cv)
            assert torch.equal(mc, mc1) and torch.equal(md, md1)
            if want_cv:
                assert torch.equal(cv, cv1)


def test_cost_volume_all_ties(ops):
    """right == 0: every disparity costs |L|_1 -> arg-min must be 0 everywhere (first index)."""
    tl = torch.randn(1, 16, 4, 37, generator=gen(3))
    tr = torch.zeros(1, 16, 4, 148)
    cv, mc, md = ops.cost_volume(nhwc(ops, tl), nhwc(ops, tr), 64, want_cv=True)
    assert torch.count_nonzero(md) == 0
    assert torch.equal(mc.cpu()[:, 0], O.l1_over_channels(tl)[:, 0])
    assert torch.equal(cv.cpu(), O.cost_volume(tl, tr, 64))


def test_cost_volume_golden(ops, golden_small):
    fx = golden_small
    d = int(fx["meta"][3])
    for k in range(5):
        tl, tr = torch.from_numpy(fx[f"tile_l{k}"]), torch.from_numpy(fx[f"tile_r{k}"])
        cv, mc, md = ops.cost_volume(nhwc(ops, tl), nhwc(ops, tr), d // (16 >> k), want_cv=True)
        assert torch.equal(cv.cpu(), torch.from_numpy(fx[f"cv{k}"]))
        assert torch.equal(md.cpu()[:, 0], torch.from_numpy(fx[f"hyp{k}"])[:, 0])


def test_tile_hyp_init(ops):
    g = gen(12)
    for cf in (16, 24, 32):
        cost = torch.rand(2, 1, 6, 10, generator=g)
        disp = torch.randint(0, 40, (2, 1, 6, 10), generator=g).float()
        feat = torch.randn(2, cf, 6, 10, generator=g)
        w = torch.randn(13, 1 + cf, 1, 1, generator=g) / 5
        b = torch.randn(13, generator=g)
        dsc = F.leaky_relu(F.conv2d(torch.cat([cost, feat], 1), w, b), 0.2)
        ref = torch.cat([disp, torch.zeros_like(disp), torch.zeros_like(disp), dsc], 1)
        for f in (nhwc(ops, feat), feat.cuda()):      # NHWC-backed and planar feature inputs
            out = ops.tile_hyp_init(cost.cuda(), disp.cuda(), f, w.cuda().contiguous(), b.cuda())
            got = back(ops, out)
            assert torch.equal(got[:, :3], ref[:, :3])
            torch.testing.assert_close(got[:, 3:], ref[:, 3:], rtol=CONV_RTOL, atol=CONV_ATOL)


# ------------------------------------------------------------------------------------------
# K3 / K4 / K5: bit-exact
# ------------------------------------------------------------------------------------------
def test_plane_upsample_bit_exact(ops):
    hyp = torch.randn(2, 16, 5, 9, generator=gen(13)) * 3
    for scale, size in ((2, 2), (1, 2), (1, 4)):
        out = ops.plane_upsample(nhwc(ops, hyp), float(scale), size)
        assert torch.equal(back(ops, out), O.plane_upsample(hyp, scale, size))


def _warp_inputs(n, c, h, w, seed, dmax):
    g = gen(seed)
    fl = torch.randn(n, c, 4 * h, 4 * w, generator=g)
    fr = torch.randn(n, c, 4 * h, 4 * w, generator=g)
    cur = torch.randn(n, 16, h, w, generator=g)
    cur[:, 0] = torch.rand(n, h, w, generator=g) * dmax - 2.0      # some samples fall off the left edge
    cur[:, 1:3] = torch.randn(n, 2, h, w, generator=g) * 0.3
    prev = torch.randn(n, 16, h // 2, w // 2, generator=g)
    prev[:, 0] = torch.rand(n, h // 2, w // 2, generator=g) * dmax / 2
    prev[:, 1:3] = torch.randn(n, 2, h // 2, w // 2, generator=g) * 0.3
    dw = torch.randn(16, 64, 1, 1, generator=g) / 8
    db = torch.randn(16, generator=g)
    return fl, fr, cur, prev, dw, db


@pytest.mark.parametrize("right_layout", ["nhwc", "planar"])
@pytest.mark.parametrize("n,c,h,w,dmax", [(1, 16, 4, 6, 10.0), (2, 24, 6, 18, 40.0), (1, 32, 2, 34, 90.0),
                                          (1, 16, 10, 16, 30.0), (1, 16, 36, 8, 20.0),
                                          # hypotheses of one CTA too far apart to stage: read in place from global memory
                                          (1, 16, 6, 80, 310.0), (1, 32, 4, 48, 180.0), (2, 24, 2, 64, 250.0)])
def test_tile_warp_cost(ops, n, c, h, w, dmax, right_layout):
    """right_layout: NHWC right features are gathered in place (codd_tile_warp_cost_nhwc); a plain contiguous NCHW
    tensor takes the planar entry point (shared-memory staged window / per-channel gathers)."""
    fl, fr, cur, prev, dw, db = _warp_inputs(n, c, h, w, c + h + w, dmax)
    fnorm = F.pixel_unshuffle(O.l1_over_channels(fl), 4)
    up_prev = O.plane_upsample(prev, 2, 2)
    raw_cur = torch.cat([fnorm, O.tile_warp_cost(cur[:, :3], fl, fr, direct=True)], 1)
    raw_prev = torch.cat([fnorm, O.tile_warp_cost(up_prev[:, :3], fl, fr, direct=True)], 1)
    dec = lambda r: F.leaky_relu(F.conv2d(r, dw, db), 0.2)
    args = (nhwc(ops, fl), nhwc(ops, fr), nhwc(ops, cur))
    # two hypothesis sets (TileUpdate)
    kw = dict(want_raw=True, force_nhwc=right_layout == "nhwc")
    aug, raw = ops.tile_warp_cost(*args, nhwc(ops, prev), dw.cuda().contiguous(), db.cuda(), **kw)
    raw = back(ops, raw)
    assert torch.equal(raw[:, :64], raw_cur), "local cost volume (current set) not bit-identical"
    assert torch.equal(raw[:, 64:], raw_prev), "local cost volume (up-sampled previous set) not bit-identical"
    aug = back(ops, aug)
    assert torch.equal(aug[:, :16], cur) and torch.equal(aug[:, 32:48], up_prev)
    torch.testing.assert_close(aug[:, 16:32], dec(raw_cur), rtol=CONV_RTOL, atol=CONV_ATOL)
    torch.testing.assert_close(aug[:, 48:], dec(raw_prev), rtol=CONV_RTOL, atol=CONV_ATOL)
    # one set (TileUpdate0)
    aug0, raw0 = ops.tile_warp_cost(*args, None, dw.cuda().contiguous(), db.cuda(), **kw)
    assert torch.equal(back(ops, raw0), raw_cur)
    aug0 = back(ops, aug0)
    assert aug0.shape[1] == 32 and torch.equal(aug0[:, :16], cur)
    torch.testing.assert_close(aug0[:, 16:], dec(raw_cur), rtol=CONV_RTOL, atol=CONV_ATOL)


@pytest.mark.parametrize("n,c,h,w,dmax", [(2, 16, 8, 40, 24), (1, 24, 6, 18, 12), (1, 32, 4, 20, 9), (1, 16, 34, 16, 60)])
def test_tile_warp_cost_integer_hypotheses(ops, n, c, h, w, dmax):
    """What the network feeds K4 as the CURRENT set: arg-min initialisations (integer disparity, zero slants,
    initialization.py:179-183).  Then ix sits within an ulp of an integer and floor() jitters by one independently per
    plane, so the planes do NOT share one 4-column window: the kernel must detect that and still be bit-exact."""
    g = gen(c * 100 + h + w)
    fl, fr, cur, prev, dw, db = _warp_inputs(n, c, h, w, c + h + w + 1, float(dmax))
    cur[:, 0] = torch.randint(0, dmax + 1, (n, h, w), generator=g).float()
    cur[:, 1:3] = 0.0
    prev[0, 0, : h // 4] = torch.randint(0, dmax // 2 + 1, (h // 4, w // 2), generator=g).float()   # some integer tiles here too
    prev[0, 1:3, : h // 4] = 0.0
    fnorm = F.pixel_unshuffle(O.l1_over_channels(fl), 4)
    up_prev = O.plane_upsample(prev, 2, 2)
    raw_cur = torch.cat([fnorm, O.tile_warp_cost(cur[:, :3], fl, fr, direct=True)], 1)
    raw_prev = torch.cat([fnorm, O.tile_warp_cost(up_prev[:, :3], fl, fr, direct=True)], 1)
    # the oracle's explicit-arithmetic form is itself bit-identical to F.grid_sample (tests/test_oracle_vs_reference.py)
    assert torch.equal(raw_cur[:, 16:], O.tile_warp_cost(cur[:, :3], fl, fr, direct=False))
    for layout in ("nhwc", "planar"):
        aug, raw = ops.tile_warp_cost(nhwc(ops, fl), nhwc(ops, fr), nhwc(ops, cur), nhwc(ops, prev), dw.cuda().contiguous(),
                                      db.cuda(), want_raw=True, force_nhwc=layout == "nhwc")
        raw = back(ops, raw)
        assert torch.equal(raw[:, :64], raw_cur), f"{layout}: integer current set not bit-identical"
        assert torch.equal(raw[:, 64:], raw_prev), f"{layout}: previous set not bit-identical"
        dec = lambda r: F.leaky_relu(F.conv2d(r, dw, db), 0.2)
        torch.testing.assert_close(back(ops, aug)[:, 16:32], dec(raw_cur), rtol=CONV_RTOL, atol=CONV_ATOL)
        torch.testing.assert_close(back(ops, aug)[:, 48:], dec(raw_prev), rtol=CONV_RTOL, atol=CONV_ATOL)


def test_tile_warp_cost_golden(ops, golden_small):
    """Against the reference's own TileWarping output (grid_sample path), finest level."""
    fx = golden_small
    sd = golden_params(fx)
    out = O.stereo_matching(sd, torch.from_numpy(fx["left"]), torch.from_numpy(fx["right"]), int(fx["meta"][3]),
                            return_all=True)
    hyp = torch.from_numpy(fx["hyp4"])
    dw, db = torch.zeros(16, 64), torch.zeros(16)
    _, raw = ops.tile_warp_cost(nhwc(ops, out["fea_l"][4]), nhwc(ops, out["fea_r"][4]), nhwc(ops, hyp), None,
                                dw.cuda(), db.cuda(), want_raw=True)
    assert torch.equal(back(ops, raw)[:, 16:], torch.from_numpy(fx["local_cv_l4"]))


def test_hyp_select_bit_exact(ops):
    g = gen(21)
    upd = torch.randn(2, 34, 7, 11, generator=g)
    upd[0, 1, :2] = upd[0, 0, :2]                      # exact confidence ties -> previous wins
    cur = torch.randn(2, 16, 7, 11, generator=g)
    upp = torch.randn(2, 16, 7, 11, generator=g)
    aug = torch.cat([cur, torch.randn(2, 16, 7, 11, generator=g), upp, torch.randn(2, 16, 7, 11, generator=g)], 1)
    ref = O.hyp_select(upd, cur, upp)[0]
    wide = ops.empty_nhwc(2, 34, 7, 11, "cuda", ld=36)
    wide.copy_(upd.cuda())
    out = ops.hyp_select(wide, nhwc(ops, aug))
    assert torch.equal(back(ops, out), ref)


# ------------------------------------------------------------------------------------------
# tensor-core 3x3 conv (tcgen05, 3xTF32)
# ------------------------------------------------------------------------------------------
TC_CASES = [(32, 32, 8, 200), (16, 16, 7, 300), (24, 24, 9, 130), (32, 16, 6, 128), (16, 3, 5, 100), (16, 1, 4, 260),
            (32, 24, 3, 60)]


@pytest.mark.parametrize("cin,cout,h,w", TC_CASES)
def test_conv3x3_tensor_core(ops, cin, cout, h, w):
    """Same bar as the fp32 CUDA-core conv (rtol/atol 2e-5): 3xTF32 must be fp32-class."""
    from codd_b200.lib import ACT_LEAKY
    g = gen(cin * 100 + cout + h)
    x = torch.randn(2, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    b = torch.randn(cout, generator=g)
    res = torch.randn(2, cout, h, w, generator=g)
    ref = F.leaky_relu(F.conv2d(x, wt, b, padding=1) + res, 0.2)
    ws = ops.pack_conv_weight_tc(wt.cuda())
    errs = {}
    for flags in (0, 1, 2, 3):
        out = ops.conv3x3_tc(nhwc(ops, x), ws, b.cuda(), cout, ACT_LEAKY, residual=nhwc(ops, res), flags=flags)
        torch.cuda.synchronize()
        errs[flags] = (back(ops, out) - ref).abs().max().item()
    print(f"conv3x3_tc cin={cin} cout={cout}: max abs err per flags {errs}")
    out = ops.conv3x3_tc(nhwc(ops, x), ws, b.cuda(), cout, ACT_LEAKY, residual=nhwc(ops, res), flags=0)
    torch.testing.assert_close(back(ops, out), ref, rtol=CONV_RTOL, atol=CONV_ATOL)


RING_CASES = [(16, 16, 1, 100), (16, 16, 2, 128), (16, 16, 3, 129), (16, 16, 40, 300), (32, 32, 5, 200), (32, 32, 23, 260),
              (24, 24, 9, 130), (32, 16, 6, 128), (16, 3, 17, 100), (16, 1, 4, 260), (32, 24, 3, 60), (16, 16, 70, 64)]


@pytest.mark.parametrize("cin,cout,h,w", RING_CASES)
def test_conv3x3_tensor_core_ring(ops, cin, cout, h, w):
    """Rolling-ring tcgen05 kernel: same bar as the other convolutions (3xTF32 must be fp32-class)."""
    from codd_b200.lib import ACT_LEAKY
    g = gen(cin * 1000 + cout * 10 + h)
    x = torch.randn(3, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    b = torch.randn(cout, generator=g)
    res = torch.randn(3, cout, h, w, generator=g)
    ref = F.leaky_relu(F.conv2d(x, wt, b, padding=1) + res, 0.2)
    out = ops.conv3x3_tc_ring(nhwc(ops, x), ops.pack_conv_weight_ring(wt.cuda()), b.cuda(), cout, ACT_LEAKY,
                              residual=nhwc(ops, res))
    torch.cuda.synchronize()
    got = back(ops, out)
    print(f"ring cin={cin} cout={cout} {h}x{w}: max abs err {(got - ref).abs().max().item():.3e}")
    torch.testing.assert_close(got, ref, rtol=CONV_RTOL, atol=CONV_ATOL)


@pytest.mark.parametrize("n,h,w,res,acts", [(2, 40, 300, True, (1, 1)), (1, 1, 100, False, (1, 1)), (3, 2, 126, True, (1, 0)),
                                              (1, 3, 127, False, (2, 1)), (2, 70, 64, True, (1, 3)), (1, 9, 253, True, (0, 1)),
                                              (1, 130, 260, False, (1, 1)), (2, 17, 1, True, (1, 1))])
def test_conv3x3_pair_fused_ring(ops, n, h, w, res, acts):
    """Two stacked 16-channel 3x3 convolutions in one rolling-ring launch (conv_merge tail / ResBlock): fp32-class vs
    torch, and BIT-IDENTICAL to two single-conv ring launches (same operand split, MMA order and epilogue arithmetic);
    strips of 126 columns (ragged widths around 126 / 252), one-row and multi-segment heights, image-border zero rows
    of the intermediate tensor."""
    g = gen(n * 1000 + h * 10 + w)
    x = torch.randn(n, 16, h, w, generator=g)
    wa = torch.randn(16, 16, 3, 3, generator=g) / 12.0
    wb = torch.randn(16, 16, 3, 3, generator=g) / 12.0
    ba, bb = torch.randn(16, generator=g), torch.randn(16, generator=g)
    act_a, act_b = acts
    xn = nhwc(ops, x)
    pa, pb = ops.pack_conv_weight_ring(wa.cuda()), ops.pack_conv_weight_ring(wb.cuda())
    assert ops.ring2_eligible(xn, 16, 16, xn if res else None) or h * w < 4096
    got = ops.conv3x3x2_tc_ring(xn, pa, ba.cuda(), act_a, pb, bb.cuda(), act_b, residual=xn if res else None)
    t = ops.conv3x3_tc_ring(xn, pa, ba.cuda(), 16, act_a)
    two = ops.conv3x3_tc_ring(t, pb, bb.cuda(), 16, act_b, residual=xn if res else None)
    torch.cuda.synchronize()

    def act(v, a):
        if a == 1:
            return F.leaky_relu(v, 0.2)
        if a == 2:
            return F.relu(v)
        if a == 3:
            return torch.cat([F.relu(v[:, :1]), v[:, 1:]], 1)
        return v
    ref = act(F.conv2d(act(F.conv2d(x, wa, ba, padding=1), act_a), wb, bb, padding=1) + (x if res else 0), act_b)
    torch.testing.assert_close(back(ops, got), ref, rtol=CONV_RTOL, atol=CONV_ATOL)
    assert torch.equal(back(ops, got), back(ops, two))


@pytest.mark.parametrize("cin,cout,n,h,w", [(16, 16, 2, 64, 256), (16, 24, 1, 130, 330), (16, 16, 1, 2, 2), (16, 32, 2, 36, 600),
                                             (16, 3, 1, 70, 258), (16, 16, 3, 128, 128), (24, 24, 2, 72, 120),
                                             (24, 32, 1, 36, 300), (32, 16, 1, 20, 260), (32, 32, 2, 18, 30)])
def test_conv4x4_stride2_tensor_core(ops, cin, cout, n, h, w):
    """conv_down first layer (backbone.py:8-14) on tcgen05: 4x4 / stride 2 / pad 1, Cin = 16, TMA boxes of same-parity
    columns, 3xTF32 — same bar as the fp32 CUDA-core convolution; ragged tile widths, single-tile and padded-filter
    (Cout = 24, 3) cases."""
    from codd_b200.lib import ACT_LEAKY
    g = gen(cout * 1000 + h + w + cin)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 4, 4, generator=g) / (16 * cin) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.leaky_relu(F.conv2d(x, wt, b, stride=2, padding=1), 0.2)
    out = ops.conv4x4s2_tc(nhwc(ops, x), ops.pack_conv_weight_tc4(wt.cuda()), b.cuda(), cout, ACT_LEAKY)
    torch.cuda.synchronize()
    got = back(ops, out)
    print(f"conv4x4s2tc cin={cin} cout={cout} {n}x{h}x{w}: max abs err {(got - ref).abs().max().item():.3e}")
    torch.testing.assert_close(got, ref, rtol=CONV_RTOL, atol=CONV_ATOL)
    # a channel slice of a wider buffer as input (ld > C)
    wide = ops.empty_nhwc(n, 48, h, w, "cuda")
    wide[:, 8:8 + cin].copy_(x.cuda())
    out2 = ops.conv4x4s2_tc(wide[:, 8:8 + cin], ops.pack_conv_weight_tc4(wt.cuda()), b.cuda(), cout, ACT_LEAKY)
    assert torch.equal(back(ops, out2), got)


@pytest.mark.parametrize("n,h,w,res", [(2, 9, 200, True), (1, 3, 128, False), (3, 36, 131, True), (1, 72, 260, True),
                                        (2, 6, 7, False)])
def test_conv3x3_ring_dilated(ops, n, h, w, res):
    """dilation 3 / pad 3, 32 -> 32 on the rolling-ring kernel (row phases as sub-images, taps 3 pixels apart): fp32-class
    against torch; one- and many-row phases, ragged strip widths, with and without the ResBlock residual."""
    from codd_b200.lib import ACT_LEAKY
    g = gen(3000 + h * w + n)
    x = torch.randn(n, 32, h, w, generator=g)
    wt = torch.randn(32, 32, 3, 3, generator=g) / (32 * 9) ** 0.5
    b = torch.randn(32, generator=g)
    r = torch.randn(n, 32, h, w, generator=g)
    ref = F.leaky_relu(F.conv2d(x, wt, b, padding=3, dilation=3) + (r if res else 0), 0.2)
    assert ops.ring_dil3_eligible(nhwc(ops, x), 32)
    out = ops.conv3x3_tc_ring(nhwc(ops, x), ops.pack_conv_weight_ring(wt.cuda()), b.cuda(), 32, ACT_LEAKY,
                              residual=nhwc(ops, r) if res else None, dil=3)
    torch.testing.assert_close(back(ops, out), ref, rtol=CONV_RTOL, atol=CONV_ATOL)


@pytest.mark.parametrize("h,w", [(9, 200), (5, 128), (20, 131), (3, 7)])
def test_conv3x3_tensor_core_dilated(ops, h, w):
    """dilation 3 / pad 3, 32 -> 32 (the dilated resblocks of tile_update4_1 / tile_update5)."""
    from codd_b200.lib import ACT_LEAKY
    g = gen(1000 + h * w)
    x = torch.randn(2, 32, h, w, generator=g)
    wt = torch.randn(32, 32, 3, 3, generator=g) / (32 * 9) ** 0.5
    b = torch.randn(32, generator=g)
    res = torch.randn(2, 32, h, w, generator=g)
    ref = F.leaky_relu(F.conv2d(x, wt, b, padding=3, dilation=3) + res, 0.2)
    assert ops.tc_eligible(32, 32, 3, 1, 3, 3, None)
    out = ops.conv3x3_tc(nhwc(ops, x), ops.pack_conv_weight_tc(wt.cuda()), b.cuda(), 32, ACT_LEAKY,
                         residual=nhwc(ops, res), dil=3)
    torch.testing.assert_close(back(ops, out), ref, rtol=CONV_RTOL, atol=CONV_ATOL)


# ------------------------------------------------------------------------------------------
# Fusion (K13)
# ------------------------------------------------------------------------------------------
def test_fusion_cue_kernels(ops):
    from oracle import fusion_oracle as FO
    outputs, state = FO.synth_fusion_inputs(2, 64, 96, seed=4)
    _, feat_warp, conf_warp, pred_warp, flow_warp = state["memory"]
    feat_curr = torch.randn(2, 32, 16, 24, generator=gen(1))
    corr_ref, fr_ref = FO.compute_input_cues(outputs["pred_disp"], pred_warp, feat_curr, feat_warp, flow_warp, conf_warp,
                                             outputs["left_feat"], outputs["right_feat"], direct=True)
    corr, disp2 = ops.fusion_cues_lowres(nhwc(ops, feat_curr), nhwc(ops, feat_warp), nhwc(ops, outputs["left_feat"]),
                                         outputs["right_feat"].cuda(), outputs["pred_disp"].cuda(), pred_warp.cuda(), 4)
    got = back(ops, corr)
    assert torch.equal(got[:, 25:31], corr_ref[:, 25:31]), "local stereo costs not bit-identical"
    torch.testing.assert_close(got[:, :25], corr_ref[:, :25], rtol=1e-5, atol=1e-5)
    d2 = back(ops, disp2)
    assert torch.equal(d2[:, 0:1], outputs["pred_disp"][..., 1::4, 1::4]) and torch.equal(d2[:, 1:2], pred_warp[..., 1::4, 1::4])
    w = torch.randn(16, 32, 1, 1, generator=gen(2)) / 6
    b = torch.randn(16, generator=gen(3))
    r16, cues = ops.fusion_forget_in(outputs["pred_disp"].cuda(), pred_warp.cuda(), flow_warp.cuda(), conf_warp.cuda(),
                                     w.cuda().contiguous(), b.cuda(), want_cues=True)
    assert torch.equal(cues.cpu(), fr_ref), "full-resolution cues not bit-identical"
    torch.testing.assert_close(back(ops, r16), F.conv2d(fr_ref, w, b), rtol=CONV_RTOL, atol=CONV_ATOL)


def test_fusion_memory_query_vs_oracle():
    import codd_b200
    from oracle import fusion_oracle as FO
    sd = FO.random_fusion_params(5)
    fus = codd_b200.MODELS.build(dict(type="Fusion", in_channels=24, fusion_channel=32,
                                      corr_cfg=dict(type="px2patch", patch_size=3)))
    fus.load_state_dict(sd, strict=True)
    fus.cuda().eval()
    for first in (True, False):
        outputs, state = FO.synth_fusion_inputs(2, 128, 192, seed=6)
        if first:
            state = {}
        ref_o = FO.memory_query(sd, {k: v.clone() for k, v in outputs.items()}, dict(state), direct=True)
        ref_s = dict(state)
        FO.memory_update(ref_o, ref_s)
        o = {k: v.cuda() for k, v in outputs.items()}
        s = {k: [t.cuda() for t in v] for k, v in state.items()}
        with torch.no_grad():
            fus.memory_query(o, s)
            fus.memory_update(o, s)
        from codd_b200 import ops as _ops
        torch.testing.assert_close(_ops.to_nchw(o["left_feat"]).cpu(), ref_o["left_feat"], rtol=1e-4, atol=1e-4)
        keys = ["pred_disp"] + ([] if first else ["fusion_weights", "reset_weights"])
        for k in keys:
            torch.testing.assert_close(o[k].cpu(), ref_o[k], rtol=1e-4, atol=1e-4)
        assert len(s["memory"]) == 3 and s["memory"][2].shape == ref_s["memory"][2].shape


# ------------------------------------------------------------------------------------------
# N1 input staging
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("h,w,to_rgb", [(540, 960, True), (375, 1242, True), (64, 64, False), (65, 130, True)])
def test_stage_images_u8(ops, h, w, to_rgb):
    import numpy as np
    from oracle import staging_oracle as SO
    rng = np.random.default_rng(h * 7 + w)
    img = rng.integers(0, 256, size=(2, h, w, 3), dtype=np.uint8)
    ref = SO.stage_images_u8(img, to_rgb=to_rgb)
    out = ops.stage_images_u8(torch.from_numpy(img).cuda(), to_rgb=to_rgb)
    assert tuple(out.shape) == ref.shape and out.shape[2] % 64 == 0 and out.shape[3] % 64 == 0
    assert torch.equal(out.cpu(), torch.from_numpy(ref)), "staging differs from the oracle"
