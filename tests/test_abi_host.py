"""CPU-side checks: the C-ABI library loads and exports everything include/codd_b200.h
declares (no compute calls), and the host-side mirror of the reference interface behaves."""
import os
import re

import pytest
import torch

import codd_b200
from codd_b200 import lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "codd_b200.h")).read()
    return re.findall(r"CODD_API\s+[\w\s\*]+?\b(codd_\w+)\s*\(", src)


def test_library_exports_every_declared_symbol():
    names = header_symbols()
    assert len(names) >= 12
    assert set(names) == set(lib.SIGNATURES), "ctypes table and header disagree"
    handle = lib.load()
    for n in names:
        assert hasattr(handle, n), n
    assert handle.codd_version() == 100
    assert handle.codd_error_string(0) == b"success"
    assert b"aligned" in handle.codd_error_string(lib.E_ALIGN)


def test_argument_errors_are_reported_without_touching_the_gpu():
    handle = lib.load()
    # null pointers / bad dims are rejected before any CUDA call
    assert handle.codd_cost_volume(None, None, 1, 1, 1, 4, None, None, None, None) == lib.E_BADARG
    assert handle.codd_plane_upsample(None, 16, 1, 1, 1, 2, 1.0, None, 16, None) == lib.E_BADARG
    assert handle.codd_cost_volume(16, 20, 1, 1, 1, 8, 16, None, None, None) == lib.E_ALIGN  # tile_r misaligned
    assert handle.codd_tile_features(16, 12, 16, 1, 8, 8, 16, 16, 16, 16, 0, 16, None) == lib.E_SHAPE  # ld < cin


def test_ops_refuse_cpu_tensors():
    x = torch.zeros(1, 16, 4, 4)
    with pytest.raises(lib.CoddError):
        ops.to_nhwc(x)
    with pytest.raises(lib.CoddError):
        ops.plane_upsample(x, 1.0, 2)


def test_registry_builds_reference_config_names():
    cfg = codd_b200.codd_stereo_config(192)
    m = codd_b200.build_estimator(cfg)
    assert type(m).__name__ == "ConsistentOnlineDynamicDepth"
    assert type(m.stereo).__name__ == "HITNetMF"
    assert m.motion is None and m.fusion is None
    assert m.stereo.tile_init.maxdisp == 192
    assert m.eval() is None  # reference quirk: train()/eval() return None
    assert not m.stereo.training
    with pytest.raises(AssertionError):
        codd_b200.build_estimator(dict(cfg, train_cfg={}), train_cfg={})


def test_state_dict_contract():
    from oracle import hitnet_oracle as O
    m = codd_b200.MODELS.build(codd_b200.hitnet_config(64))
    sd = m.state_dict()
    ref = O.random_hitnet_params(1)
    assert set(sd) == set(ref)
    assert all(sd[k].shape == ref[k].shape for k in sd)
    m.load_state_dict(ref, strict=True)
    # checkpoints trained with a loss carry extra tensors; strict=False must tolerate them
    ref2 = dict(ref)
    ref2["loss.convx.weight"] = torch.zeros(1, 1, 9, 9)
    missing, unexpected = m.load_state_dict(ref2, strict=False)
    assert not missing and unexpected == ["loss.convx.weight"]


def test_weight_packing_layouts():
    w = torch.arange(2 * 3 * 4 * 5, dtype=torch.float32).view(2, 3, 4, 5)   # [Cout,Cin,KH,KW]
    p = ops.pack_conv_weight(w)
    assert p.shape == (4, 5, 3, 2)
    assert p[1, 2, 0, 1] == w[1, 0, 1, 2]
    wt = torch.arange(3 * 2 * 2 * 2, dtype=torch.float32).view(3, 2, 2, 2)  # [Cin,Cout,2,2]
    q = ops.pack_deconv_weight(wt)
    assert q.shape == (2, 2, 3, 2) and q[1, 0, 2, 1] == wt[2, 1, 1, 0]


def test_nhwc_view_helpers():
    t = ops.empty_nhwc(2, 16, 3, 5, "cpu", ld=64)
    assert t.shape == (2, 16, 3, 5) and ops.ld_of(t) == 64
    assert ops.ld_of(t[:, 4:8]) == 64 and ops.ld_of(t[1:]) == 64
    c = torch.zeros(2, 16, 3, 5)
    with pytest.raises(lib.CoddError):
        ops.ld_of(c)
    assert ops.ld_of(c.contiguous(memory_format=torch.channels_last)) == 16


def test_training_paths_fail_loudly():
    m = codd_b200.build_estimator(codd_b200.codd_stereo_config(64))
    with pytest.raises(NotImplementedError):
        m(img=[torch.zeros(1, 1, 3, 64, 64)], img_metas=[[{}]], return_loss=True)


def test_full_codd_config_builds_reference_parameter_tree():
    """stereo + motion + fusion through the registry: names / shapes of SURVEY.md Appendix B."""
    m = codd_b200.build_estimator(codd_b200.codd_full_config(192, iters=16))
    assert type(m.motion).__name__ == "Motion" and type(m.motion.raft3d).__name__ == "RAFT3D"
    assert type(m.motion.raft3d.cnet[0]).__name__ == "HRNet" and type(m.fusion).__name__ == "Fusion"
    sd = m.state_dict()
    want = {
        "motion.raft3d.fnet.conv1.weight": (64, 3, 7, 7),
        "motion.raft3d.fnet.layer2.0.downsample.0.weight": (96, 64, 1, 1),
        "motion.raft3d.cnet.0.conv1.weight": (64, 3, 3, 3),
        "motion.raft3d.cnet.0.layer1.0.downsample.0.weight": (256, 64, 1, 1),
        "motion.raft3d.cnet.0.transition1.1.0.0.weight": (36, 256, 3, 3),
        "motion.raft3d.cnet.0.stage4.1.fuse_layers.3.0.2.0.weight": (144, 18, 3, 3),
        "motion.raft3d.cnet.0.stage3.2.fuse_layers.0.2.0.weight": (18, 72, 1, 1),
        "motion.raft3d.cnet.1.convs.0.weight": (512, 270, 1, 1),
        "motion.raft3d.update_block.gru.convq2.weight": (128, 128, 3, 3),
        "motion.raft3d.update_block.corr_enc.0.weight": (256, 196, 3, 3),
        "motion.raft3d.update_block.flow_enc.0.weight": (128, 9, 7, 7),
        "motion.raft3d.update_block.mask.2.weight": (576, 256, 1, 1),
        "fusion.motion_conv.0.weight": (30, 64, 7, 7),
        "stereo.tile_update.tile_update4_1.resblocks.1.0.conv1.0.0.weight": (32, 32, 3, 3),
    }
    for k, shape in want.items():
        assert k in sd and tuple(sd[k].shape) == shape, k
    assert "motion.raft3d.cnet.1.convs.0.bias" not in sd
    n_update = sum(v.numel() for k, v in sd.items() if k.startswith("motion.raft3d.update_block."))
    n_fnet = sum(v.numel() for k, v in sd.items() if k.startswith("motion.raft3d.fnet."))
    assert abs(n_update / 1e6 - 3.47) < 0.01 and abs(n_fnet / 1e6 - 1.05) < 0.01
    assert m.eval() is None and not m.motion.training


def test_ring_weight_packing_reconstructs_weights():
    """ops.pack_conv_weight_ring (host side of codd_conv3x3_tc_ring): pass A rows per ky = [fp16(w) | fp16(2^10 (w - fp16(w)))],
    pass B = [0 | fp16(w)]; hi + lo / 1024 reproduces w to 2^-22 relative, padding rows / columns are zero."""
    import torch
    from codd_b200 import ops
    g = torch.Generator().manual_seed(3)
    for cout, cin in [(16, 16), (24, 24), (32, 32), (1, 16), (16, 32)]:
        w = torch.randn(cout, cin, 3, 3, generator=g) * 0.3
        buf = ops.pack_conv_weight_ring(w)
        kc = 16 if cin <= 16 else 32
        npad = 16 if cout <= 16 else 32
        h = buf.view(torch.float16).view(2, 3, 3, 2, npad, kc).float()      # [pass][kx][ky][hi|lo][cout][cin]
        pa, pb = h[0], h[1]
        rec = (pa[:, :, 0] + pa[:, :, 1] / 1024.0)[:, :, :cout, :cin]       # [kx][ky][cout][cin]
        ref = w.permute(3, 2, 0, 1)
        assert torch.allclose(rec, ref, rtol=0, atol=float(ref.abs().max()) * 2.0 ** -21)
        assert torch.equal(pb[:, :, 1], pa[:, :, 0]) and not pb[:, :, 0].any()
        assert not pa[:, :, :, cout:].any() and not pa[:, :, :, :, cin:].any()


def test_tc4_weight_packing_reconstructs_weights():
    """ops.pack_conv_weight_tc4 (host side of codd_conv4x4s2_tc / codd_tile_features_tc): per tap NP rows of fp16(w) followed
    by NP rows of fp16(2^10 (w - fp16(w))); hi + lo / 1024 reproduces w to 2^-21 relative, padding is zero."""
    import torch
    from codd_b200 import ops
    g = torch.Generator().manual_seed(5)
    for cout, cin in [(16, 16), (24, 16), (24, 24), (32, 24), (16, 32), (3, 16)]:
        w = torch.randn(cout, cin, 4, 4, generator=g) * 0.3
        buf = ops.pack_conv_weight_tc4(w)
        kc = 16 if cin <= 16 else 32
        npad = 16 if cout <= 16 else 32
        h = buf.view(torch.float16).view(16, 2, npad, kc).float()          # [tap][hi|lo][cout][cin]
        rec = (h[:, 0] + h[:, 1] / 1024.0)[:, :cout, :cin]
        ref = w.permute(2, 3, 0, 1).reshape(16, cout, cin)
        assert torch.allclose(rec, ref, rtol=0, atol=float(ref.abs().max()) * 2.0 ** -21)
        assert not h[:, :, cout:].any() and not h[:, :, :, cin:].any()
