"""N3: the on-GPU evaluation kernels (through the C ABI) against the oracle restatement of codd.py:462-515."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as M

pytestmark = pytest.mark.gpu

_spec = importlib.util.spec_from_file_location("_metrics_cases", os.path.join(os.path.dirname(__file__), "test_metrics_oracle.py"))
_cases = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_cases)

RANGE = (0.0, 192.0)


@pytest.fixture(scope="module")
def ops():
    from codd_b200 import ops as o
    return o


def dev(c):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in c.items() if v.dtype != bool}


@pytest.mark.parametrize("seed,kitti,n,h,w", [(0, False, 1, 23, 37), (1, False, 2, 40, 64), (2, True, 1, 23, 37),
                                               (3, False, 1, 135, 240)])
def test_frame_metrics_vs_oracle(ops, seed, kitti, n, h, w):
    c = _cases.make_case(seed, n=n, h=h, w=w, kitti=kitti)
    d = dev(c)
    # the prediction lives in a padded buffer, as the network output does
    pad = torch.full((n, 1, h + 9, w + 27), -7.0, device="cuda")
    pad[:, :, :h, :w] = d["pred"]
    pad_prev = torch.full((n, 1, h + 9, w + 27), -7.0, device="cuda")
    pad_prev[:, :, :h, :w] = d["pred_prev"]
    acc = torch.zeros(16, dtype=torch.float64, device="cuda")
    mask = torch.empty((n, 1, h, w), dtype=torch.uint8, device="cuda")
    ops.disp_metrics(pad, d["gt"], RANGE, acc[0:4], seg=d["seg"], mask_out=mask)
    mask_prev = torch.from_numpy(c["mask_prev"].astype(np.uint8)).cuda()
    ops.temporal_metrics(d["flow"], d["gt"], pad, d["gt_prev"], pad_prev, mask_prev, RANGE, acc[4:13], seg=d["seg"],
                         gt_pos_count=acc[3:4])
    a = acc.cpu().numpy()
    mask_o = M.valid_mask(c["gt"], RANGE, seg=c["seg"])
    do = M.disp_metrics(c["pred"], c["gt"], mask_o)
    to = M.temporal_metrics(c["flow"], c["gt"], c["pred"], c["seg"], c["gt_prev"], c["pred_prev"], c["mask_prev"], RANGE)
    assert np.array_equal(mask.cpu().numpy().astype(bool), mask_o)                 # bit-exact masks
    assert a[0] == do["n"] and a[3] == int((c["gt"] > 0).sum())                    # exact counts
    if do["n"]:
        assert a[1] / a[0] == pytest.approx(do["epe"], rel=1e-12)
        assert a[2] / a[0] == do["th3"]
    t = a[4:13]
    assert t[8] == n * h * w and t[7] / t[8] == pytest.approx(to["flow_mag"], rel=1e-12)
    assert bool(t[5] > 0 and t[6] > 0) == to["updated"]
    if to["updated"]:
        assert t[0] == to["n"]
        assert t[1] / t[0] == pytest.approx(to["tepe"], rel=1e-12)
        assert t[2] / t[0] == pytest.approx(to["tepe_rel"], rel=1e-12)
        assert t[3] / t[0] == to["th1_tepe_rel"] and t[4] / t[0] == to["th3_tepe"]  # thresholded counts are exact


def test_sequence_metrics_matches_per_frame_oracle(ops):
    """SequenceMetrics over 4 frames == the reference's AverageMeter bookkeeping over the oracle's per-frame values."""
    from codd_b200.metrics import SequenceMetrics
    frames = [_cases.make_case(20 + i, n=1, h=31, w=45) for i in range(4)]
    sm = SequenceMetrics(RANGE, max_frames=8)
    exp = {k: [] for k in ("epe", "th3", "tepe", "tepe_rel", "th1_tepe_rel", "th3_tepe", "flow_mag")}
    prev = None
    for f in frames:
        d = dev(f)
        sm.update(d["pred"], d["gt"], gt_flow=d["flow"], seg=d["seg"])
        mask = M.valid_mask(f["gt"], RANGE, seg=f["seg"])
        do = M.disp_metrics(f["pred"], f["gt"], mask)
        if do["n"]:
            exp["epe"].append(do["epe"]); exp["th3"].append(do["th3"])
        if prev is not None:
            to = M.temporal_metrics(prev["flow"], f["gt"], f["pred"], f["seg"], prev["gt"], prev["pred"], prev["mask"], RANGE)
            exp["flow_mag"].append(to["flow_mag"])
            if to["updated"]:
                for k in ("tepe", "tepe_rel", "th1_tepe_rel", "th3_tepe"):
                    exp[k].append(to[k])
        prev = dict(flow=f["flow"], gt=f["gt"], pred=f["pred"], mask=mask)
    got = sm.collect()
    for k, v in exp.items():
        assert got[k] == pytest.approx(float(np.mean(v)), rel=1e-10), k


def make_motion_case(seed, n, h, w):
    g = np.random.default_rng(seed)
    q = g.normal(0, 0.03, (n, h + 5, w + 11, 4)).astype(np.float32)
    q[..., 3] += 1.0
    q /= np.linalg.norm(q, axis=-1, keepdims=True)
    Ts = np.concatenate([g.normal(0, 0.05, (n, h + 5, w + 11, 3)).astype(np.float32), q], -1)
    pred_prev = g.uniform(-2, 120, (n, 1, h, w)).astype(np.float32)
    pred_prev[g.random(pred_prev.shape) < 0.05] = 0.0
    intr = np.tile(np.array([[450.0, 460.0, w / 2.0, h / 2.0]], np.float32), (n, 1))
    flow = g.normal(0, 2, (n, 2, h, w)).astype(np.float32)
    flow[g.random(flow.shape) < 0.02] = 300.0
    dc = g.normal(0, 1.0, (n, 1, h, w)).astype(np.float32)
    dc[g.random(dc.shape) < 0.05] = 210.0                      # BF_DEFAULT marks invalid disparity change
    gt_prev = g.uniform(-3, 220, (n, 1, h, w)).astype(np.float32)
    seg = (g.random((n, 1, h, w)) > 0.1).astype(np.float32)
    occ = g.random((n, 1, h, w)) < 0.1
    return dict(Ts=Ts, pred_prev=pred_prev, intr=intr, flow=flow, dc=dc, gt_prev=gt_prev, seg=seg, occ=occ)


@pytest.mark.parametrize("seed,n,h,w,use_occ", [(0, 1, 24, 40, True), (1, 2, 33, 57, False)])
def test_sceneflow_metrics_vs_oracle(ops, seed, n, h, w, use_occ):
    c = make_motion_case(seed, n, h, w)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in c.items()}
    acc = torch.zeros(5, dtype=torch.float64, device="cuda")
    ops.sceneflow_metrics(d["Ts"][:, :h, :w], d["pred_prev"], d["intr"], d["flow"], d["dc"], d["gt_prev"], RANGE, acc,
                          seg=d["seg"], flow_occ=d["occ"] if use_occ else None)
    a = acc.cpu().numpy()
    o = M.sceneflow_metrics(c["Ts"][:, :h, :w], c["pred_prev"], c["intr"], c["flow"], c["dc"], c["gt_prev"], RANGE,
                            seg=c["seg"], flow_occ=c["occ"] if use_occ else None)
    assert a[0] == o["n"] and o["n"] > 0                                      # the mask is exact
    assert a[1] == pytest.approx(o["sum_sf"], rel=1e-4) and a[2] == pytest.approx(o["sum_of"], rel=1e-4)
    assert abs(a[3] - o["n1_sf"]) <= 2 and abs(a[4] - o["n1_of"]) <= 2        # fp32 contraction at the 1-px boundary


def test_sequence_metrics_motion_block(ops):
    """SequenceMetrics.update(..., Ts=, intrinsics=, gt_disp_change=) accumulates the running sums of codd.py:567-575."""
    from codd_b200.metrics import SequenceMetrics
    h, w = 31, 45
    frames = [_cases.make_case(40 + i, n=1, h=h, w=w) for i in range(3)]
    motions = [make_motion_case(50 + i, 1, h, w) for i in range(3)]
    sm = SequenceMetrics(RANGE, max_frames=8)
    exp = np.zeros(5)
    prev = None
    for f, mo in zip(frames, motions):
        d = dev(f)
        md = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in mo.items()}
        sm.update(d["pred"], d["gt"], gt_flow=d["flow"], seg=d["seg"], Ts=md["Ts"][:, :h, :w], intrinsics=md["intr"],
                  gt_disp_change=md["dc"], gt_flow_occ_prev=md["occ"])
        if prev is not None:
            o = M.sceneflow_metrics(mo["Ts"][:, :h, :w], prev["pred"], mo["intr"], prev["flow"], mo["dc"], prev["gt"], RANGE,
                                    seg=f["seg"], flow_occ=mo["occ"])
            exp += np.array([o["n"], o["sum_sf"], o["sum_of"], o["n1_sf"], o["n1_of"]], np.float64)
        prev = dict(pred=f["pred"], gt=f["gt"], flow=f["flow"])
    got = sm.collect()
    assert got["count"] == exp[0] and exp[0] > 0
    assert got["epe2d_scene_flow"] == pytest.approx(exp[1], rel=1e-4)
    assert got["epe2d_optical_flow"] == pytest.approx(exp[2], rel=1e-4)
    assert abs(got["1px_scene_flow"] - exp[3]) <= 3 and abs(got["1px_optical_flow"] - exp[4]) <= 3


def test_gt_disp_change_vs_oracle(ops):
    c = _cases.make_case(5, n=2, h=33, w=47)
    occ = np.random.default_rng(9).random(c["gt"].shape) < 0.2
    d = dev(c)
    change, warped = ops.gt_disp_change(d["flow"], d["gt"], d["gt_prev"], torch.from_numpy(occ).cuda())
    o_change, o_warped = M.gt_disp_change(occ, c["gt_prev"], c["gt"], c["flow"])
    assert np.array_equal(change.cpu().numpy(), o_change) and np.array_equal(warped.cpu().numpy(), o_warped)
