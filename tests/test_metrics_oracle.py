"""N3 (on-GPU metrics) oracle pinned against the reference's own functions (utils/misc.py, utils/warp.py,
utils/metric.py) driven in the order model/codd.py:462-515 calls them.  CPU only; skipped when /root/reference
is absent (the golden fixture test below still runs)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as M
from oracle import ref_loader

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "metrics_golden.json")


def make_case(seed, n=1, h=23, w=37, kitti=False):
    g = np.random.default_rng(seed)
    gt = g.uniform(-5, 230, (n, 1, h, w)).astype(np.float32)
    gt[g.random((n, 1, h, w)) < 0.15] = 0.0
    pred = (gt + g.normal(0, 2.5, gt.shape)).astype(np.float32)
    gt_prev = g.uniform(0.5, 200, (n, 1, h, w)).astype(np.float32)
    pred_prev = (gt_prev + g.normal(0, 2.5, gt.shape)).astype(np.float32)
    # flows: sub-pixel, exact half-integers (nearest ties), out of the image, and huge magnitudes (> BF)
    flow = g.normal(0, 3, (n, 2, h, w)).astype(np.float32)
    ties = g.random((n, 2, h, w)) < 0.2
    flow[ties] = (np.round(flow[ties]) + 0.5).astype(np.float32)
    flow[g.random((n, 2, h, w)) < 0.02] = 400.0
    seg = (g.random((n, 1, h, w)) > 0.1).astype(np.float32)
    if kitti:
        gt[:] = 0.0
    mask_prev = M.valid_mask(gt_prev, (0.0, 192.0), seg=seg)
    return dict(gt=gt, pred=pred, gt_prev=gt_prev, pred_prev=pred_prev, flow=flow, seg=seg, mask_prev=mask_prev)


def reference_metrics(U, c, disp_range=(0.0, 192.0)):
    """The reference call sequence of codd.py:462-515 on torch tensors."""
    t = {k: torch.from_numpy(v) for k, v in c.items()}
    meta = {"disp_range": disp_range}
    gt, pred, flow, seg = t["gt"], t["pred"], t["flow"], t["seg"]
    mask_disp = U.compute_valid_mask(gt, meta, gt_semantic_seg=seg)
    out = {}
    if mask_disp.any():
        out["epe"] = torch.mean(torch.abs(pred[mask_disp] - gt[mask_disp])).item()
        out["th3"] = U.thres_metric(pred, gt, mask_disp, 3.0).item()
    if torch.any(gt > 0.0):
        mask = U.compute_valid_mask(gt, meta, gt_flow_prev=flow, gt_semantic_seg=seg)
    else:
        mask = U.compute_valid_mask(torch.ones_like(gt) * U.BF_DEFAULT / 2.0, meta, gt_flow_prev=flow,
                                    gt_semantic_seg=seg)
    to_warp = torch.cat([gt, pred, mask.float()], dim=1)
    to_warp, valid = U.flow_warp(to_warp, flow, padding_mode="zeros", mode="nearest")
    w_gt, w_pred, w_mask = torch.unbind(to_warp, dim=1)
    w_gt, w_pred = w_gt.unsqueeze(1), w_pred.unsqueeze(1)
    mask_curr = (valid.squeeze()[0] & w_mask.bool() & mask)
    mask_prev = t["mask_prev"]
    if mask_prev.any() and mask_curr.any():
        a, r = U.t_epe_metric(w_pred, w_gt, t["pred_prev"], t["gt_prev"], mask_prev, mask_curr)
        out.update(tepe=a.mean().item(), tepe_rel=r.mean().item(), th1=(r > 1.0).float().mean().item(),
                   th3_tepe=(a > 3.0).float().mean().item(), n=int(a.numel()))
    out["flow_mag"] = torch.sum(flow ** 2, dim=1).sqrt().squeeze().mean().item()
    out["mask_disp"] = mask_disp.numpy()
    out["warp"] = (to_warp.numpy(), valid.numpy())
    return out


def oracle_metrics(c, disp_range=(0.0, 192.0)):
    mask_disp = M.valid_mask(c["gt"], disp_range, seg=c["seg"])
    d = M.disp_metrics(c["pred"], c["gt"], mask_disp)
    t = M.temporal_metrics(c["flow"], c["gt"], c["pred"], c["seg"], c["gt_prev"], c["pred_prev"], c["mask_prev"],
                           disp_range)
    return mask_disp, d, t


@pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference")
@pytest.mark.parametrize("seed,kitti", [(0, False), (1, False), (2, True), (3, False)])
def test_oracle_matches_reference(seed, kitti):
    U = ref_loader.load().utils
    c = make_case(seed, kitti=kitti)
    ref = reference_metrics(U, c)
    mask_disp, d, t = oracle_metrics(c)
    assert np.array_equal(mask_disp, ref["mask_disp"])
    # the nearest warp itself is bit-exact (index arithmetic)
    gt, pred = c["gt"], c["pred"]
    if (gt > 0).any():
        mask = M.valid_mask(gt, (0.0, 192.0), seg=c["seg"], flow_prev=c["flow"])
    else:
        mask = M.valid_mask(np.full_like(gt, M.BF_DEFAULT / np.float32(2)), (0.0, 192.0), seg=c["seg"],
                            flow_prev=c["flow"])
    warped, valid = M.flow_warp_nearest(np.concatenate([gt, pred, mask.astype(np.float32)], 1), c["flow"])
    assert np.array_equal(warped, ref["warp"][0]) and np.array_equal(valid, ref["warp"][1])
    if d["n"]:
        assert d["epe"] == pytest.approx(ref["epe"], rel=1e-5)
        assert d["th3"] == pytest.approx(ref["th3"], rel=1e-6)
    else:
        assert "epe" not in ref
    assert t["flow_mag"] == pytest.approx(ref["flow_mag"], rel=1e-5)
    assert t["updated"] == ("tepe" in ref)
    if t["updated"]:
        assert t["n"] == ref["n"]
        assert t["tepe"] == pytest.approx(ref["tepe"], rel=1e-5)
        assert t["tepe_rel"] == pytest.approx(ref["tepe_rel"], rel=1e-5)
        assert t["th1_tepe_rel"] == pytest.approx(ref["th1"], rel=1e-6)
        assert t["th3_tepe"] == pytest.approx(ref["th3_tepe"], rel=1e-6)


def test_oracle_matches_golden_fixture():
    """Values produced by the REFERENCE functions in the build container (tests/golden/gen_metrics_golden.py)."""
    gold = json.load(open(GOLDEN))
    for entry in gold:
        c = make_case(entry["seed"], kitti=entry["kitti"])
        _, d, t = oracle_metrics(c)
        r = entry["ref"]
        if "epe" in r:
            assert d["epe"] == pytest.approx(r["epe"], rel=1e-5) and d["th3"] == pytest.approx(r["th3"], rel=1e-6)
        assert t["flow_mag"] == pytest.approx(r["flow_mag"], rel=1e-5)
        assert t["updated"] == ("tepe" in r)
        if t["updated"]:
            assert t["n"] == r["n"]
            for k_o, k_r in [("tepe", "tepe"), ("tepe_rel", "tepe_rel"), ("th1_tepe_rel", "th1"), ("th3_tepe", "th3_tepe")]:
                assert t[k_o] == pytest.approx(r[k_r], rel=1e-5)


@pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference")
def test_gt_disp_change_matches_reference():
    U = ref_loader.load().utils
    c = make_case(5, n=2)
    occ = np.random.default_rng(9).random(c["gt"].shape) < 0.2
    ref_change, ref_warp = U.compute_gt_disp_change(torch.from_numpy(occ), torch.from_numpy(c["gt_prev"]),
                                                    torch.from_numpy(c["gt"]), torch.from_numpy(c["flow"]))
    change, warped = M.gt_disp_change(occ, c["gt_prev"], c["gt"], c["flow"])
    assert np.array_equal(change, ref_change.numpy()) and np.array_equal(warped, ref_warp.numpy())
