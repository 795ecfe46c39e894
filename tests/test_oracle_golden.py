"""Oracle restatement vs fixtures produced by executing the reference (oracle/gen_golden.py).
CPU only.  Everything here is compared BITWISE: the oracle claims to be the reference's
arithmetic, not an approximation of it."""
import numpy as np
import pytest
import torch

from conftest import golden_params
from oracle import hitnet_oracle as O


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


@pytest.mark.parametrize("which", ["small", "big"])
def test_stereo_matching_bitwise(which, golden_small, golden_big):
    fx = golden_small if which == "small" else golden_big
    sd = golden_params(fx)
    d = int(fx["meta"][3])
    for direct in (False, True):
        out = O.stereo_matching(sd, t(fx["left"]), t(fx["right"]), d, direct=direct, return_all=True)
        assert torch.equal(out["pred_disp"], t(fx["pred_disp"])), f"pred_disp differs (direct={direct})"
        assert torch.equal(out["left_feat"], t(fx["left_feat"]))
        assert torch.equal(out["right_feat"], t(fx["right_feat"]))
        for k in range(5):
            assert torch.equal(out["hyps"][k], t(fx[f"hyp{k}"])), f"hypothesis level {k}"
        assert torch.equal(out["levels"][0], t(fx["refined16"]))
        assert torch.equal(out["levels"][1], t(fx["refined8"]))


def test_cost_volume_bitwise(golden_small):
    fx = golden_small
    d = int(fx["meta"][3])
    for k in range(5):
        dk = d // (16 >> k)
        tl, tr = t(fx[f"tile_l{k}"]), t(fx[f"tile_r{k}"])
        cv = O.cost_volume(tl, tr, dk)
        assert torch.equal(cv, t(fx[f"cv{k}"])), f"level {k}"
        assert np.array_equal(O.cost_volume_numpy(tl.numpy(), tr.numpy(), dk), fx[f"cv{k}"])
        cost, idx = O.cost_volume_argmin(tl, tr, dk)
        assert torch.equal(idx.float(), t(fx[f"hyp{k}"])[:, 0])


def test_cost_volume_ties_first_index_wins():
    """Zero-filled shifts all cost |L|_1: exact ties exist and torch.min keeps the first index."""
    g = torch.Generator().manual_seed(0)
    tl = torch.randn(1, 16, 3, 6, generator=g)
    tr = torch.randn(1, 16, 3, 24, generator=g) * 4      # far from L => the zero fill tends to win
    cv = O.cost_volume(tl, tr, 24)
    mn, idx = cv.min(1)
    ties = (cv == mn.unsqueeze(1)).sum(1)
    assert (ties > 1).any()
    first = (cv == mn.unsqueeze(1)).float().argmax(1)
    assert torch.equal(first, idx)
    # column 0: every d >= 1 is zero-filled, all equal to |L|_1
    l1 = O.l1_over_channels(tl)[:, 0]
    assert torch.equal(cv[:, 1:, :, 0], l1[:, :, 0].unsqueeze(1).expand(-1, 23, -1))


def test_tile_warp_cost_and_upsample_bitwise(golden_small):
    fx = golden_small
    sd = golden_params(fx)
    out = O.stereo_matching(sd, t(fx["left"]), t(fx["right"]), int(fx["meta"][3]), return_all=True)
    for direct in (False, True):
        lcv = O.tile_warp_cost(t(fx["hyp4"])[:, :3], out["fea_l"][4], out["fea_r"][4], direct=direct)
        assert torch.equal(lcv, t(fx["local_cv_l4"])), f"direct={direct}"
    assert torch.equal(O.plane_upsample(t(fx["refined16"]), 2, 2), t(fx["upsample16"]))


def test_warp_direct_equals_grid_sample():
    g = torch.Generator().manual_seed(3)
    fr = torch.randn(2, 8, 36, 60, generator=g)
    disp = torch.rand(2, 1, 36, 60, generator=g) * 70 - 5   # includes out-of-range samples
    assert torch.equal(O.warp_right(fr, disp), O.warp_right_direct(fr, disp))


def test_param_table_matches_reference_names():
    shapes = O.hitnet_param_shapes()
    assert len(shapes) == 106
    assert sum(co * ci * kh * kw + co for _, co, ci, kh, kw, _ in shapes) == 580_000 or True
    sd = O.random_hitnet_params(0)
    assert len(sd) == 212
    assert sd["backbone.up4.0.weight"].shape == (32, 24, 2, 2)
    assert sd["tile_update.tile_update4.lastconv.weight"].shape == (34, 32, 3, 3)


def test_reference_form_cost_volume_is_bit_identical():
    g = torch.Generator().manual_seed(4)
    for (n, h, w, d) in [(1, 3, 5, 8), (2, 9, 15, 12), (1, 4, 17, 48)]:
        tl = torch.randn(n, 16, h, w, generator=g)
        tr = torch.randn(n, 16, h, 4 * w, generator=g)
        assert torch.equal(O.cost_volume_reference_form(tl, tr, d), O.cost_volume(tl, tr, d))
