"""RAFT3D networks + Motion module on the GPU against oracle/raft3d_oracle.py.
Float tolerances: single layers 2e-5 relative (fp32 FMA order), whole networks 1e-3 of the output scale."""
import pytest
import torch
import torch.nn.functional as F

from oracle import motion_oracle as M
from oracle import raft3d_oracle as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available()
    from codd_b200 import ops as _ops
    return _ops


def g(seed):
    return torch.Generator().manual_seed(seed)


def close(got, want, tol=1e-3):
    scale = want.abs().max().clamp(min=1e-6)
    err = (got - want).abs().max() / scale
    assert err <= tol, f"max err {err:.3e} of scale {scale:.3e}"


@pytest.mark.parametrize("cin,cout,k,s,p,d", [
    (64, 96, 3, 2, 1, 1), (3, 64, 7, 2, 3, 1), (64, 96, 1, 2, 0, 1), (128, 128, 3, 1, 4, 4),
    (196, 256, 3, 1, 1, 1), (128, 1024, 3, 1, 1, 1), (256, 576, 1, 1, 0, 1), (270, 512, 1, 1, 0, 1), (9, 128, 7, 1, 3, 1)])
def test_generic_conv_geometries(ops, cin, cout, k, s, p, d):
    x = torch.randn(2, cin, 17, 22, generator=g(1))
    w = torch.randn(cout, cin, k, k, generator=g(2)) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g(3))
    want = F.relu(F.conv2d(x, w, b, stride=s, padding=p, dilation=d))
    got = ops.conv2d(ops.to_nhwc(x.cuda()), ops.pack_conv_weight(w.cuda()), b.cuda(), cout, k, s, p, d, ops.ACT_RELU)
    assert got.shape == want.shape
    close(ops.to_nchw(got).cpu(), want, 2e-5)


def test_instance_norm_resize_eltwise(ops):
    x = torch.randn(2, 96, 20, 31, generator=g(4)) * 3 + 1
    res = torch.randn(2, 96, 20, 31, generator=g(5))
    xc, rc = ops.to_nhwc(x.cuda()), ops.to_nhwc(res.cuda())
    close(ops.to_nchw(ops.instance_norm(xc, relu=False)).cpu(), F.instance_norm(x), 1e-5)
    close(ops.to_nchw(ops.instance_norm(xc, relu=True, residual=rc)).cpu(), F.relu(res + F.relu(F.instance_norm(x))), 1e-5)
    for align, size in ((False, (40, 62)), (True, (33, 47)), (True, (10, 16))):
        want = F.interpolate(x, size=size, mode="bilinear", align_corners=align)
        close(ops.to_nchw(ops.resize_bilinear(xc, size, align)).cpu(), want, 1e-5)
    base = torch.randn(2, 96, 40, 62, generator=g(6))
    want = F.relu(base + F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False))
    close(ops.to_nchw(ops.resize_bilinear(xc, (40, 62), False, base=ops.to_nhwc(base.cuda()), relu=True)).cpu(), want, 1e-5)
    z = torch.rand(2, 96, 20, 31, generator=g(7))
    close(ops.to_nchw(ops.eltwise(ops.EW_GRU, ops.to_nhwc(z.cuda()), xc, rc)).cpu(), (1 - z) * x + z * res, 1e-6)
    close(ops.to_nchw(ops.eltwise(ops.EW_MUL, xc, rc)).cpu(), x * res, 1e-6)
    close(ops.to_nchw(ops.eltwise(ops.EW_ACT, xc, act=ops.ACT_TANH)).cpu(), torch.tanh(x), 1e-5)
    close(ops.to_nchw(ops.eltwise(ops.EW_ADD_ACT, xc, rc, act=ops.ACT_RELU)).cpu(), F.relu(x + res), 1e-6)
    disp = torch.rand(2, 24, 40, generator=g(8)) * 100
    disp[0, :3] = 0.0
    want = torch.clip(210.0 / (disp + 1e-5), max=210.0, min=0)
    torch.testing.assert_close(ops.disp_to_depth(disp.cuda(), 210.0).cpu(), want, rtol=1e-6, atol=0)
    t = torch.randn(2, 24, 40, 7, generator=g(9))
    assert torch.equal(ops.subsample(t.cuda(), 1, 4).cpu(), t[:, 1::4, 1::4])
    assert torch.equal(ops.subsample(t.cuda(), 3, 8, recip=True).cpu(), 1.0 / t[:, 3::8, 3::8])


@pytest.fixture(scope="module")
def raft():
    """(state_dict, CUDA module) of a randomly initialised RAFT3D."""
    import codd_b200
    from codd_b200.motion import RAFT3D
    sd = R.random_raft3d_params(7)
    net = RAFT3D(cnet_cfg=dict(type="HRNet", norm_eval=True, extra=codd_b200.HRNET_W18_SMALL))
    net.load_state_dict(sd)
    return sd, net.cuda().eval()


def test_feature_and_context_networks(ops, raft):
    sd, net = raft
    img = torch.randn(2, 3, 128, 192, generator=g(10))
    with torch.no_grad():
        f = net.fnet(img.cuda())
        feats = net.cnet[0](img.cuda())
        c = net.context(img.cuda())
    close(ops.to_nchw(f).cpu(), R.basic_encoder(sd, "fnet.", img), 1e-3)
    want = R.hrnet(sd, "cnet.0.", img)
    assert [tuple(t.shape) for t in feats] == [tuple(t.shape) for t in want]
    for a, b in zip(feats, want):
        close(ops.to_nchw(a).cpu(), b, 1e-3)
    close(ops.to_nchw(c).cpu(), R.context_net(sd, "cnet.", img), 1e-3)


def test_update_block(ops, raft):
    sd, net = raft
    n, h, w = 2, 16, 24
    netf, inp = torch.randn(n, 128, h, w, generator=g(11)).tanh(), torch.randn(n, 384, h, w, generator=g(12)).relu()
    corr = torch.randn(n, 196, h, w, generator=g(13))
    flow, dz, twist = torch.randn(n, h, w, 2, generator=g(14)) * 3, torch.randn(n, h, w, 1, generator=g(15)), \
        torch.randn(n, h, w, 6, generator=g(16))
    want = R.update_block(sd, "update_block.", netf, inp, corr, flow, dz, twist)
    with torch.no_grad():
        got = net.update_block(netf.cuda(), inp.cuda(), corr.cuda(), flow.cuda(), dz.cuda(), twist.cuda())
    for a, b in zip(got, want):
        close(ops.to_nchw(a).cpu(), b, 1e-3)


def _two_frames(n, h, w, seed):
    """A moving textured scene: frame pair, disparities and intrinsics."""
    from oracle import hitnet_oracle as O
    img0, _ = O.synth_pair(n, h, w, 32, seed=seed, kind="S")
    img1 = torch.roll(img0, shifts=(1, 2), dims=(2, 3))
    yy, xx = torch.meshgrid(torch.arange(h).float(), torch.arange(w).float(), indexing="ij")
    disp0 = (8.0 + 6.0 * torch.sin(xx / 23.0) * torch.cos(yy / 17.0)).expand(n, 1, h, w).contiguous()
    disp1 = torch.roll(disp0, shifts=(1, 2), dims=(2, 3)) + 0.3
    return img0, img1, disp0, disp1, [float(w), float(w), w / 2.0, h / 2.0]


def test_motion_forward_two_frames(ops, raft):
    """Motion.forward on frame 0 (priming) and frame 1 (RAFT3D loop + both splat warps) vs the oracle."""
    import codd_b200
    from codd_b200.motion import Motion
    sd, net = raft
    n, h, w, iters = 1, 128, 192, 3
    img0, img1, disp0, disp1, intr = _two_frames(n, h, w, 31)
    feat0 = torch.randn(n, 32, h // 4, w // 4, generator=g(17))
    metas = [dict(intrinsics=intr)]

    mot = Motion(raft3d=dict(type="RAFT3D", cnet_cfg=dict(type="HRNet", norm_eval=True, extra=codd_b200.HRNET_W18_SMALL)),
                 iters=iters)
    mot.raft3d.load_state_dict(sd)
    mot = mot.cuda().eval()
    msd = {"raft3d." + k: v for k, v in sd.items()}

    st_ref, st = {}, {}
    out_ref = dict(left_img=img0, pred_disp=disp0)
    out = dict(left_img=img0.cuda(), pred_disp=disp0.cuda())
    with torch.no_grad():
        R.motion_forward(msd, "", st_ref, out_ref, intr, iters)
        mot(st, out, img_metas=metas)
        close(ops.to_nchw(st["raft_feat"]).cpu(), st_ref["raft_feat"], 1e-3)
        close(ops.to_nchw(st["raft_netinp"]).cpu(), st_ref["raft_netinp"], 1e-3)
        st_ref["memory"] = [img0, feat0, disp0.squeeze(1)]
        st["memory"] = [img0.cuda(), ops.to_nhwc(feat0.cuda()), disp0.squeeze(1).cuda()]
        out_ref = dict(left_img=img1, pred_disp=disp1)
        out = dict(left_img=img1.cuda(), pred_disp=disp1.cuda())
        R.motion_forward(msd, "", st_ref, out_ref, intr, iters)
        mot(st, out, img_metas=metas)
    # the SE3 field and what is derived from it
    Tg, Tr = out["Ts"].cpu(), out_ref["Ts"]
    assert Tg.shape == Tr.shape == (n, h, w, 7)
    close(Tg, Tr, 2e-3)
    close(out["flow2d_est_induced"].cpu(), out_ref["flow2d_est_induced"], 5e-3)
    close(ops.to_nchw(ops.to_nhwc(out["weight"])).cpu(), out_ref["weight"], 2e-3)
    # warped memory: splatting is discontinuous in the point positions -> compare robustly
    names = ["img_warp", "feat_warp", "confidence_warp", "disp_warp", "flow_warp"]
    for name, a, b in zip(names, st["memory"], st_ref["memory"]):
        a = ops.to_nchw(ops.to_nhwc(a)).cpu() if a.dim() == 4 else a.cpu()
        assert a.shape == b.shape, name
        bad = ((a - b).abs() > 1e-2 * b.abs().clamp(min=1.0)).float().mean().item()
        assert bad < 0.01, f"{name}: {bad * 100:.2f}% of elements differ"


@pytest.mark.parametrize("cin,cout,k,p,d,n,h,w", [
    (128, 256, 3, 1, 1, 2, 24, 40), (196, 256, 3, 1, 1, 1, 33, 47), (128, 128, 3, 4, 4, 2, 24, 40),
    (256, 384, 1, 0, 1, 2, 24, 40), (128, 1024, 3, 1, 1, 1, 20, 60), (12, 128, 7, 3, 1, 2, 24, 40),
    (256, 576, 1, 0, 1, 1, 30, 36), (256, 32, 1, 0, 1, 2, 24, 40)])
def test_conv_as_tcgen05_gemm(ops, cin, cout, k, p, d, n, h, w):
    """im2col + tcgen05 GEMM (3xTF32) against fp32 F.conv2d: same bar as every other convolution."""
    x = torch.randn(n, cin, h, w, generator=g(21))
    wt = torch.randn(cout, cin, k, k, generator=g(22)) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g(23))
    res = torch.randn(n, cout, h, w, generator=g(24))
    want = F.relu(F.conv2d(x, wt, b, padding=p, dilation=d) + res)
    assert ops.gemm_eligible(n, h, w, cin, cout, k, 1, None)
    got = ops.conv2d_gemm(ops.to_nhwc(x.cuda()), ops.pack_conv_weight_gemm(wt.cuda()), b.cuda(), cout, k, p, d, ops.ACT_RELU,
                          residual=ops.to_nhwc(res.cuda()))
    close(ops.to_nchw(got).cpu(), want, 2e-5)
