"""Motion / RAFT3D non-convolutional kernels (K9-K12) against oracle/motion_oracle.py.
The oracle itself is pinned only in part (see its header): these are consistency tests of the CUDA
path with the CPU restatement, float tolerance 1e-4 relative (transcendentals, fp32 vs fp64 solve)."""
import pytest
import torch

from oracle import motion_oracle as M

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available()
    from codd_b200 import ops as _ops
    return _ops


def g(seed):
    return torch.Generator().manual_seed(seed)


def rand_Ts(n, h, w, seed, scale=0.05):
    return M.se3_exp(torch.randn(n, h, w, 6, generator=g(seed)) * scale)


def test_motion_info(ops):
    n, h, w = 2, 9, 15
    intr = torch.tensor([[60., 62., 7., 4.], [58., 60., 7.5, 4.5]])
    d1 = 1.0 + torch.rand(n, h, w, generator=g(1)) * 4
    d2 = 1.0 + torch.rand(n, h, w, generator=g(2)) * 4
    Ts = rand_Ts(n, h, w, 3)
    xyz_ref, info_ref = M.motion_info(Ts, d1, d2, intr)
    xyz, info = ops.raft_motion_info(Ts.cuda(), d1.cuda(), (1.0 / d2).cuda(), intr.cuda())
    torch.testing.assert_close(xyz.cpu(), xyz_ref, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ops.to_nchw(info).cpu().permute(0, 2, 3, 1), info_ref, rtol=1e-4, atol=2e-4)


def test_corr_lookup(ops):
    n, c, h, w = 2, 128, 8, 16
    f1 = torch.randn(n, c, h, w, generator=g(4))
    f2 = torch.randn(n, c, h, w, generator=g(5))
    coords = torch.rand(n, h, w, 3, generator=g(6)) * torch.tensor([w + 2., h + 2., 1.]) - 1.0
    ref = M.corr_lookup(M.all_pairs_correlation(f1, f2, 3), coords[..., :2].permute(0, 3, 1, 2), radius=3)
    pyr = ops.corr_pyramid(ops.to_nhwc(f2.cuda()), 3)
    out = ops.corr_lookup(ops.to_nhwc(f1.cuda()), pyr, coords.cuda(), radius=3)
    torch.testing.assert_close(ops.to_nchw(out).cpu(), ref, rtol=1e-4, atol=1e-4)


def test_gn_step(ops):
    n, h, w = 1, 6, 10
    intr = torch.tensor([[40., 42., 5., 3.]])
    depth = 2.0 + torch.rand(n, h, w, generator=g(7))
    Ts = rand_Ts(n, h, w, 8, 0.02)
    ae = torch.randn(n, 32, h, w, generator=g(9))
    target = M.project(M.se3_act(rand_Ts(n, h, w, 10, 0.03), M.inv_project(depth, intr)), intr).permute(0, 3, 1, 2).contiguous()
    weight = torch.rand(n, 3, h, w, generator=g(11))
    for radius in (2, 32):
        ref = M.gn_step(Ts, ae, target, weight, depth, intr, radius=radius)
        out = ops.se3_gn_step(Ts.cuda(), ops.to_nhwc(ae.cuda()), ops.to_nhwc(target.cuda()), ops.to_nhwc(weight.cuda()),
                              depth.cuda(), intr.cuda(), radius=radius)
        torch.testing.assert_close(out.cpu(), ref, rtol=2e-4, atol=2e-4)


def test_cvx_upsample_and_se3_flow(ops):
    n, h, w = 2, 5, 7
    intr = torch.tensor([[300., 310., 28., 20.], [290., 300., 27., 19.]])
    data = torch.randn(n, h, w, 3, generator=g(12))
    mask = torch.randn(n, 576, h, w, generator=g(13))
    out = ops.cvx_upsample(data.cuda(), ops.to_nhwc(mask.cuda()))
    torch.testing.assert_close(out.cpu(), M.cvx_upsample(data, mask), rtol=1e-5, atol=1e-5)
    Ts = rand_Ts(n, h, w, 14)
    depth = 1.0 + torch.rand(n, 8 * h, 8 * w, generator=g(15)) * 5
    up, flow = ops.se3_upsample_flow(Ts.cuda(), ops.to_nhwc(mask.cuda()), depth.cuda(), intr.cuda())
    up_ref = M.upsample_se3(Ts, mask)
    torch.testing.assert_close(up.cpu(), up_ref, rtol=1e-4, atol=1e-5)
    flow_ref, _ = M.induced_flow(up_ref, depth, intr)
    torch.testing.assert_close(flow.cpu(), flow_ref, rtol=1e-3, atol=2e-3)


def test_splat_warp(ops):
    n, c, h, w = 1, 5, 12, 16
    intr = torch.tensor([[40., 40., 8., 6.]])
    depth = 2.0 + torch.rand(n, h, w, generator=g(16)) * 2
    feat = torch.randn(n, c, h, w, generator=g(17))
    Ts = rand_Ts(n, h, w, 18, 0.03)
    for radius in (2.0, 4.0):
        ref, zref = M.splat_warp(Ts, depth, feat, intr, radius)
        out, zbuf, disp = ops.splat_warp(Ts.cuda(), depth.cuda(), intr.cuda(), ops.to_nhwc(feat.cuda()), radius, bf=210.0,
                                         want_disp=True)
        torch.testing.assert_close(ops.to_nchw(out).cpu(), ref, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(zbuf.cpu(), zref, rtol=1e-5, atol=1e-5)
        dref = 210.0 / (zref + 1e-5)
        dref[dref > w] = 0.0
        torch.testing.assert_close(disp.cpu(), dref, rtol=1e-4, atol=1e-4)


def test_splat_warp_crowded_pixels_exact_and_deterministic(ops):
    """More than 32 points on one pixel (low-res feature warp at radius 4 after a converging motion): the
    renderer must keep exactly the 8 nearest in z — like pytorch3d's points_per_pixel — whatever order the atomics land
    in, and two runs must agree bit for bit."""
    n, c, h, w = 1, 4, 16, 16
    intr = torch.tensor([[30., 30., 8., 8.]])
    depth = 2.0 + torch.rand(n, h, w, generator=g(31)) * 3
    feat = torch.randn(n, c, h, w, generator=g(32))
    Ts = rand_Ts(n, h, w, 33, 0.01)
    # converge: pull every point toward the optical axis so that the central pixels collect most of the image
    Ts[..., 0:2] = -0.7 * (torch.stack(torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")[::-1], -1).float()
                           - 8.0).unsqueeze(0) * depth.unsqueeze(-1) / 30.0
    radius = 6.0
    ref, zref = M.splat_warp(Ts, depth, feat, intr, radius)
    # crowding check on the oracle's own candidate count
    X = M.se3_act(Ts, M.inv_project(depth, intr)).reshape(-1, 3)
    u, v = 30 * X[:, 0] / X[:, 2] + 8, 30 * X[:, 1] / X[:, 2] + 8
    s, r = 2.0 / 16, radius / h
    most = max(int(((((u - (px + .5)) * s) ** 2 + ((v - (py + .5)) * s) ** 2) < r * r).sum()) for py in range(6, 10) for px in range(6, 10))
    assert most > 32, most
    outs = []
    for _ in range(3):
        out, zbuf, _ = ops.splat_warp(Ts.cuda(), depth.cuda(), intr.cuda(), ops.to_nhwc(feat.cuda()), radius)
        outs.append((ops.to_nchw(out).cpu(), zbuf.cpu()))
    assert all(torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1]) for o in outs[1:]), "not deterministic"
    torch.testing.assert_close(outs[0][1], zref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(outs[0][0], ref, rtol=1e-4, atol=1e-4)
