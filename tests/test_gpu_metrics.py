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
