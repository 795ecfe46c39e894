"""Pins the oracle against the UNMODIFIED reference files, where they are mounted
(/root/reference exists in the build container, not on the GPU box -> skipped there)."""
import warnings

import pytest
import torch

from oracle import hitnet_oracle as O
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference not mounted")


@pytest.fixture(scope="module")
def ref_model():
    warnings.filterwarnings("ignore")
    m = ref_loader.build_hitnet(64)
    m.load_state_dict(O.random_hitnet_params(21))
    return m


@pytest.mark.parametrize("kind,shape", [("S", (1, 128, 192)), ("G", (2, 128, 128)), ("U", (1, 192, 128))])
def test_end_to_end_bitwise(ref_model, kind, shape):
    n, h, w = shape
    left, right = O.synth_pair(n, h, w, 64, seed=99, kind=kind)
    with torch.no_grad():
        ref = ref_model.stereo_matching(left, right)
    sd = {k: v.detach() for k, v in ref_model.state_dict().items()}
    for direct in (False, True):
        out = O.stereo_matching(sd, left, right, 64, direct=direct)
        for key in ("pred_disp", "left_feat", "right_feat"):
            assert torch.equal(out[key], ref[key]), (key, direct)


def test_calc_init_disp_bitwise_ragged():
    ns = ref_loader.load()
    g = torch.Generator().manual_seed(5)
    for (n, h, w, d) in [(1, 3, 5, 8), (2, 9, 15, 12), (1, 4, 7, 32), (1, 2, 40, 48)]:
        tl = torch.randn(n, 16, h, w, generator=g)
        tr = torch.randn(n, 16, h, 4 * w, generator=g)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = ns.initialization.calc_init_disp(tl, tr, d)
        assert torch.equal(O.cost_volume(tl, tr, d), ref)


def test_upsample_and_warp_bitwise():
    ns = ref_loader.load()
    g = torch.Generator().manual_seed(6)
    hyp = torch.randn(2, 16, 5, 7, generator=g)
    assert torch.equal(O.plane_upsample(hyp, 2, 2), ns.propagation.upsample(hyp, 2))
    assert torch.equal(O.plane_upsample(hyp, 1, 2), ns.propagation.upsample(hyp, 1))
    assert torch.equal(O.plane_upsample(hyp, 16, 64), ns.propagation.upsample(hyp, 16, 64))
    fr = torch.randn(2, 24, 20, 28, generator=g)
    disp = torch.rand(2, 1, 20, 28, generator=g) * 40 - 4
    ref = ns.propagation.warp(fr, disp)
    assert torch.equal(O.warp_right(fr, disp), ref)
    assert torch.equal(O.warp_right_direct(fr, disp), ref)


def test_same_seed_same_weights_as_reference():
    """The drop-in modules consume the RNG in the reference's construction order."""
    import codd_b200
    torch.manual_seed(123)
    mine = codd_b200.MODELS.build(codd_b200.hitnet_config(64))
    torch.manual_seed(123)
    ref = ref_loader.build_hitnet(64)
    a, b = mine.state_dict(), ref.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_fusion_oracle_vs_reference():
    """Fusion.memory_query / memory_update (model/fusion/fusion.py:357-410) through the shim."""
    from oracle import fusion_oracle as FO
    ns = ref_loader.load()
    assert ns.fusion is not None, getattr(ns, "fusion_error", None)
    ref = ns.fusion.Fusion(in_channels=24, fusion_channel=32, corr_cfg=dict(type="px2patch", patch_size=3))
    sd = FO.random_fusion_params(3)
    ref.load_state_dict(sd, strict=True)
    ref.eval()
    assert set(sd) == set(ref.state_dict())
    for first in (True, False):
        outputs, state = FO.synth_fusion_inputs(2, 64, 96, seed=9)
        if first:
            state = {}
        o_ref = {k: v.clone() for k, v in outputs.items()}
        s_ref = {k: [t.clone() for t in v] for k, v in state.items()}
        with torch.no_grad():
            ref.memory_query(o_ref, s_ref)
            ref.memory_update(o_ref, s_ref)
            o_or = FO.memory_query(sd, {k: v.clone() for k, v in outputs.items()}, state)
            s_or = dict(state)
            FO.memory_update(o_or, s_or)
        keys = ["pred_disp", "left_feat"] + ([] if first else ["fusion_weights", "reset_weights"])
        for k in keys:
            torch.testing.assert_close(o_or[k], o_ref[k], rtol=1e-5, atol=1e-5)
        for a, b in zip(s_or["memory"], s_ref["memory"]):
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)


# ----------------------------------------------------------------------------------------------
# RAFT3D networks (oracle/raft3d_oracle.py) against the reference's own classes
# ----------------------------------------------------------------------------------------------
def _ref_raft3d():
    import importlib
    ref_loader.load()
    return (importlib.import_module("model.motion.raft3d.raft3d"),
            importlib.import_module("model.motion.raft3d.blocks.extractor"))


def test_basic_encoder_matches_reference():
    from oracle import raft3d_oracle as R
    _, extractor = _ref_raft3d()
    torch.manual_seed(3)
    ref = extractor.BasicEncoder(output_dim=128, norm_fn="instance").eval()
    sd = {"fnet." + k: v.detach() for k, v in ref.state_dict().items()}
    x = torch.randn(2, 3, 64, 96)
    with torch.no_grad():
        want = ref(x)
    torch.testing.assert_close(R.basic_encoder(sd, "fnet.", x), want, rtol=1e-5, atol=1e-5)
    # the codd_b200 container exposes the same parameter names / shapes
    from codd_b200.motion.extractor import BasicEncoder
    mine = BasicEncoder(output_dim=128, norm_fn="instance").state_dict()
    assert {k: tuple(v.shape) for k, v in mine.items()} == {k: tuple(v.shape) for k, v in ref.state_dict().items()}


def test_update_block_and_resize_concat_match_reference():
    from oracle import raft3d_oracle as R
    raft3d, _ = _ref_raft3d()
    torch.manual_seed(4)
    ref = raft3d.BasicUpdateBlock(hidden_dim=128).eval()
    sd = {"update_block." + k: v.detach() for k, v in ref.state_dict().items()}
    n, h, w = 1, 6, 9
    net, inp, corr = torch.randn(n, 128, h, w).tanh(), torch.randn(n, 384, h, w).relu(), torch.randn(n, 196, h, w)
    flow, dz, twist = torch.randn(n, h, w, 2) * 3, torch.randn(n, h, w, 1), torch.randn(n, h, w, 6)
    with torch.no_grad():
        want = ref(net, inp, corr, flow, dz, twist)          # the call site's argument order (raft3d.py:238-240)
    got = R.update_block(sd, "update_block.", net, inp, corr, flow, dz, twist)
    for g_, w_ in zip(got, want):
        torch.testing.assert_close(g_, w_, rtol=1e-5, atol=1e-5)
    from codd_b200.motion.raft3d import BasicUpdateBlock
    mine = BasicUpdateBlock(hidden_dim=128).state_dict()
    assert {k: tuple(v.shape) for k, v in mine.items()} == {k: tuple(v.shape) for k, v in ref.state_dict().items()}

    rcc = raft3d.ResizeConcatConv([18, 36, 72, 144], 512).eval()
    sd = {"cnet.1." + k: v.detach() for k, v in rcc.state_dict().items()}
    feats = [torch.randn(1, c, 32 >> i, 48 >> i) for i, c in enumerate([18, 36, 72, 144])]
    with torch.no_grad():
        want = rcc(feats)
    torch.testing.assert_close(R.resize_concat(sd, "cnet.1.", feats), want, rtol=1e-5, atol=1e-5)
