"""End-to-end parity of the drop-in HITNetMF (CUDA path through the C ABI) against the CPU
oracle and the reference-generated golden fixtures, plus size-independent properties at the
BASELINE.json size (576x960, D=192).

Tolerance: north_star asks for <= 1e-3 relative on disparity.  The discrete selections in the
network (arg-min tile init, arg-max hypothesis select) can flip on ulp-level differences of the
convolutions and move isolated tiles; the tests therefore assert the 1e-3 bound on >= 99.5 % of
the pixels (mixed abs/rel: |a-b| <= 1e-3 * max(1, |b|)), report the flip statistics, and assert
the intermediate arg-min indices separately."""
import numpy as np
import pytest
import torch

from conftest import golden_params
from oracle import hitnet_oracle as O

pytestmark = pytest.mark.gpu


def build(max_disp, sd):
    import codd_b200
    m = codd_b200.MODELS.build(codd_b200.hitnet_config(max_disp))
    m.load_state_dict(sd, strict=True)
    m.cuda().eval()
    return m


def frac_within(a, b, tol=1e-3):
    err = (a - b).abs()
    ok = err <= tol * b.abs().clamp(min=1.0)
    return ok.float().mean().item(), err.max().item()


def gpu_argmin(stereo, left, right):
    """The CUDA path's arg-min initialisation pyramid (coarse->fine), as CPU tensors [N,h,w]."""
    from codd_b200 import ops
    with torch.no_grad():
        fl, fr = stereo.backbone.forward_pair(left.cuda(), right.cuda())
        _, hyps = stereo.tile_init(fl, fr)
    return [ops.to_nchw(h).cpu()[:, 0] for h in hyps]


def parity_modulo_near_ties(stereo, sd, left, right, d, pred, ref_pred, tag=""):
    """Fraction of pixels within 1e-3 of the oracle; if a near-tie of the oracle's cost volume was resolved the other
    way by the CUDA path (certified by oracle.stereo_matching_given_argmin), against the oracle's result for those
    choices."""
    frac, mx = frac_within(pred, ref_pred)
    print(f"{tag} pred_disp: {frac*100:.3f}% within 1e-3, max abs err {mx:.3e}")
    if frac < 0.995:
        ref2 = O.stereo_matching_given_argmin(sd, left.cpu(), right.cpu(), d, gpu_argmin(stereo, left, right))
        print(f"{tag} {ref2['flips']} arg-min near-tie flip(s), {ref2['uncertified']} uncertified")
        assert ref2["flips"] > 0 and ref2["uncertified"] == 0
        frac, mx = frac_within(pred, ref2["pred_disp"][:, :, :pred.shape[2], :pred.shape[3]])
        print(f"{tag} pred_disp given the certified choices: {frac*100:.3f}% within 1e-3, max abs err {mx:.3e}")
    return frac


@pytest.mark.parametrize("which", ["small", "big"])
def test_stereo_matching_vs_golden(which, golden_small, golden_big):
    from codd_b200 import ops
    fx = golden_small if which == "small" else golden_big
    sd = golden_params(fx)
    d = int(fx["meta"][3])
    m = build(d, sd)
    left, right = torch.from_numpy(fx["left"]).cuda(), torch.from_numpy(fx["right"]).cuda()
    with torch.no_grad():
        out = m.stereo_matching(left, right)
        fl, fr = m.backbone.forward_pair(left, right)
        cvs, hyps = m.tile_init(fl, fr)
    pred = out["pred_disp"].cpu()
    assert pred.shape == fx["pred_disp"].shape and pred.is_contiguous()
    # features: pure conv stack -> tolerance
    torch.testing.assert_close(ops.to_nchw(out["left_feat"]).cpu(), torch.from_numpy(fx["left_feat"]),
                               rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ops.to_nchw(out["right_feat"]).cpu(), torch.from_numpy(fx["right_feat"]),
                               rtol=1e-4, atol=1e-4)
    # arg-min tile initialisation per level (end-to-end: inputs differ by conv rounding)
    for k in range(5):
        got = ops.to_nchw(hyps[k]).cpu()[:, 0]
        ref = torch.from_numpy(fx[f"hyp{k}"])[:, 0]
        agree = (got == ref).float().mean().item()
        differ = int((got != ref).sum())
        print(f"[{which}] level {k}: arg-min agreement {agree:.5f} ({differ} of {ref.numel()} differ)")
        # end to end, the inputs of K1 differ from the reference's by conv rounding (~1e-7), which can flip near-ties
        # (measured: at most ONE tile per level on these fixtures, two in total at 576x960, profiles/parity_r02.json);
        # bit-exactness on identical inputs is asserted in test_gpu_ops.py, the flips themselves are certified below
        assert differ <= 2
    frac = parity_modulo_near_ties(m, sd, left, right, d, pred, torch.from_numpy(fx["pred_disp"]), f"[{which}]")
    assert frac >= 0.995


def test_stereo_matching_vs_oracle_structured():
    """Set S (textured pair with a smooth disparity field), 256x320, D=64, batch 2."""
    sd = O.random_hitnet_params(42)
    left, right = O.synth_pair(2, 256, 320, 64, seed=77, kind="S")
    ref = O.stereo_matching(sd, left, right, 64, direct=True)
    m = build(64, sd)
    with torch.no_grad():
        out = m.stereo_matching(left.cuda(), right.cuda())
    frac = parity_modulo_near_ties(m, sd, left, right, 64, out["pred_disp"].cpu(), ref["pred_disp"], "structured")
    assert frac >= 0.995


def test_codd_top_level_api():
    """model(return_loss=False, rescale=True, evaluate=False, img=[..], img_metas=[[..]], r_img=[..])"""
    import codd_b200
    torch.manual_seed(0)
    model = codd_b200.build_estimator(codd_b200.codd_stereo_config(64)).cuda()
    model.eval()
    left, right = O.synth_pair(2, 128, 192, 64, seed=5, kind="S")
    img = torch.stack([left, left.flip(0)], 1).cuda()       # [B, MF=2, 3, H, W]
    r_img = torch.stack([right, right.flip(0)], 1).cuda()
    metas = [[dict(min_disp=1, max_disp=64, ori_shape=(120, 190), img_shape=(120, 190))]]
    res = model(return_loss=False, rescale=True, evaluate=False, img=[img], img_metas=metas, r_img=[r_img])
    assert isinstance(res, list) and res[0].shape == (2, 2, 120, 190)
    sd = {k[len("stereo."):]: v.detach().cpu() for k, v in model.state_dict().items()}
    ref = O.stereo_matching(sd, left, right, 64, direct=True)["pred_disp"][:, :, :120, :190]
    frac = parity_modulo_near_ties(model.stereo, sd, left, right, 64, res[0][:, 0:1].cpu(), ref, "api")
    assert frac >= 0.995


def test_full_size_properties():
    """576x960, D=192 (BASELINE.json configs[1] geometry, batch 2 to bound test time):
    properties that need no CPU oracle at this size."""
    import codd_b200
    from codd_b200 import ops
    torch.manual_seed(0)
    m = codd_b200.MODELS.build(codd_b200.hitnet_config(192)).cuda().eval()
    left, right = O.synth_pair(2, 576, 960, 192, seed=1234, kind="S")
    left, right = left.cuda(), right.cuda()
    with torch.no_grad():
        fl, fr = m.backbone.forward_pair(left, right)
        tiles = m.tile_init.tile_features(fl, fr)
        for k in range(5):
            tl, tr = tiles[k]
            dk = 192 // (16 >> k)
            cv, mc, md = ops.cost_volume(tl, tr, dk, want_cv=True)
            _, mc2, md2 = ops.cost_volume(tl, tr, dk, want_cv=False)
            # fused arg-min == materialising variant, and both consistent with the volume:
            assert torch.equal(mc, mc2) and torch.equal(md, md2)
            assert torch.equal(cv.min(1, keepdim=True)[0], mc)
            idx = md.long()
            assert torch.equal(cv.gather(1, idx), mc)
            # first-index rule: no strictly earlier disparity reaches the minimum
            earlier = torch.arange(dk, device="cuda").view(1, -1, 1, 1) < idx
            assert not ((cv <= mc) & earlier).any()
            # zero-filled shifts (4j - d < -3 ... < 0) all equal |L|_1
            l1 = ops.to_nchw(tl).abs()
            acc = l1[:, 0]
            for c in range(1, 16):
                acc = acc + l1[:, c]
            assert torch.equal(cv[:, dk - 1, :, 0], acc[:, :, 0])
        out = m.stereo_matching(left, right)
        out2 = m.stereo_matching(left, right)
    pred = out["pred_disp"]
    assert pred.shape == (2, 1, 576, 960) and torch.isfinite(pred).all() and (pred >= 0).all()
    assert torch.equal(pred, out2["pred_disp"]), "forward is not deterministic"
    # batch independence: sample 1 alone gives the same disparity as inside the batch
    with torch.no_grad():
        solo = m.stereo_matching(left[1:], right[1:])["pred_disp"]
    assert torch.equal(solo, pred[1:])


def test_kitti_shape_properties():
    """384x1280 (KITTI 1242x375 padded: BASELINE.json configs[3] geometry), D=192, batch 2: finite, non-negative,
    deterministic, batch-independent, and cropped correctly by the top-level API."""
    import codd_b200
    torch.manual_seed(1)
    model = codd_b200.build_estimator(codd_b200.codd_stereo_config(192)).cuda()
    model.eval()
    left, right = O.synth_pair(2, 384, 1280, 192, seed=7, kind="S")
    img, r_img = left.unsqueeze(1).cuda(), right.unsqueeze(1).cuda()
    metas = [[dict(min_disp=1, max_disp=192, ori_shape=(375, 1242), img_shape=(375, 1242))]]
    res = model(return_loss=False, rescale=True, evaluate=False, img=[img], img_metas=metas, r_img=[r_img])[0]
    assert res.shape == (2, 1, 375, 1242) and torch.isfinite(res).all() and (res >= 0).all()
    res2 = model(return_loss=False, rescale=True, evaluate=False, img=[img], img_metas=metas, r_img=[r_img])[0]
    assert torch.equal(res, res2)
    solo = model(return_loss=False, rescale=True, evaluate=False, img=[img[1:]], img_metas=metas, r_img=[r_img[1:]])[0]
    assert torch.equal(solo, res[1:])
