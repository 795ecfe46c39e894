"""evaluate=True bookkeeping of the drop-in estimator (codd.py:313-355, 519-540) on CPU: the model's per-frame
forward and the GPU meters are replaced by recorders, so only the ground-truth plumbing is under test."""
import torch

import codd_b200
from codd_b200 import metrics as metrics_mod


class _Recorder:
    calls = []

    def __init__(self, disp_range, max_frames=0, device=None):
        _Recorder.calls = []
        self.disp_range = disp_range

    def update(self, pred, gt, **kw):
        _Recorder.calls.append(dict(pred=pred, gt=gt, **kw))

    def collect(self):
        return {"epe": 1.5, "count": 7.0}


def test_evaluate_true_feeds_the_meters(monkeypatch):
    monkeypatch.setattr(metrics_mod, "SequenceMetrics", _Recorder)
    model = codd_b200.build_estimator(codd_b200.codd_stereo_config(64))
    model.eval()
    B, MF, H, W, h, w = 1, 3, 64, 128, 50, 100
    g = torch.Generator().manual_seed(0)
    preds = [torch.rand(B, 1, H, W, generator=g) for _ in range(MF)]
    Ts = [torch.rand(B, H, W, 7, generator=g) for _ in range(MF)]
    it = iter(range(MF))

    def fake_frame(l, r, metas, state):
        i = next(it)
        return {"pred_disp": preds[i], "Ts": Ts[i]}

    monkeypatch.setattr(model, "consistent_online_depth_estimation", fake_frame)
    img = torch.zeros(B, MF, 3, H, W)
    gt = torch.rand(B, MF, 1, H, W, generator=g) * 60
    flow = torch.randn(B, MF, 2, H, W, generator=g)
    dc = torch.randn(B, MF, 1, H, W, generator=g)
    occ = (torch.rand(B, MF, 1, H, W, generator=g) > 0.8).float()
    metas = [[dict(img_shape=(h, w, 3), disp_range=(0.0, 64.0), intrinsics=[100.0, 110.0, 50.0, 25.0])]]
    res = model(return_loss=False, rescale=True, evaluate=True, img=[img], img_metas=metas, r_img=[img],
                gt_disp=[gt], gt_flow=[flow], gt_disp_change=[dc], gt_disp_occ=[occ])
    assert isinstance(res, list) and torch.equal(res[0]["epe"], torch.tensor([1.5])) and res[0]["count"].item() == 7.0
    calls = _Recorder.calls
    assert len(calls) == MF
    for i, c in enumerate(calls):
        assert c["pred"] is preds[i]                                              # uncropped; the meters crop to gt's size
        assert torch.equal(c["gt"], gt[:, i, :, :h, :w]) and torch.equal(c["gt_flow"], flow[:, i, :, :h, :w])
        assert torch.equal(c["seg"], (occ[:, i] <= 0)[:, :, :h, :w].float())      # True = not occluded (codd.py:350-353)
        if i == 0:
            assert "Ts" not in c
        else:                                                                      # provided disp change: entry [-2]
            assert torch.equal(c["gt_disp_change"], dc[:, i - 1, :, :h, :w])
            assert torch.equal(c["Ts"], Ts[i][:, :h, :w]) and c["intrinsics"].shape == (B, 4)
            assert "gt_flow_occ_prev" not in c


def test_evaluate_true_with_disp2_or_flow_occ(monkeypatch):
    """gt_disp2 without gt_disp_change: change = disp2 - disp with BF marking invalid pixels (codd.py:343-349), previous
    entry [-2] used; gt_disp_change together with gt_flow_occ: the current entry [-1] and the previous occlusion mask
    [-2] are used (codd.py:521-527)."""
    monkeypatch.setattr(metrics_mod, "SequenceMetrics", _Recorder)
    model = codd_b200.build_estimator(codd_b200.codd_stereo_config(64))
    model.eval()
    B, MF, H, W = 1, 3, 64, 64
    g = torch.Generator().manual_seed(1)
    monkeypatch.setattr(model, "consistent_online_depth_estimation",
                        lambda l, r, m, s: {"pred_disp": torch.ones(B, 1, H, W), "Ts": torch.zeros(B, H, W, 7)})
    img = torch.zeros(B, MF, 3, H, W)
    gt = torch.rand(B, MF, 1, H, W, generator=g) * 60 - 5
    gt2 = torch.rand(B, MF, 1, H, W, generator=g) * 60 - 5
    dc = torch.randn(B, MF, 1, H, W, generator=g)
    focc = (torch.rand(B, MF, 1, H, W, generator=g) > 0.7).float()
    zero_flow = torch.zeros(B, MF, 2, H, W)
    metas = [[dict(img_shape=(H, W, 3), disp_range=(0.0, 64.0), intrinsics=[100.0, 100.0, 32.0, 32.0])]]

    model(return_loss=False, evaluate=True, img=[img], img_metas=metas, r_img=[img], gt_disp=[gt], gt_flow=[zero_flow],
          gt_disp2=[gt2])
    c = _Recorder.calls[2]

    def change(i):
        e = gt2[:, i] - gt[:, i]
        e[gt2[:, i] <= 0] = 210.0
        e[gt[:, i] <= 0] = 210.0
        return e

    assert torch.equal(c["gt_disp_change"], change(1)) and "gt_flow_occ_prev" not in c
    assert torch.equal(c["gt_disp2"], gt2[:, 2])

    model(return_loss=False, evaluate=True, img=[img], img_metas=metas, r_img=[img], gt_disp=[gt], gt_flow=[zero_flow],
          gt_disp_change=[dc], gt_flow_occ=[focc])
    c = _Recorder.calls[2]
    assert torch.equal(c["gt_disp_change"], dc[:, 2]) and torch.equal(c["gt_flow_occ_prev"], focc[:, 1] > 0)


def test_evaluate_true_derives_disp_change_from_flow(monkeypatch):
    """gt_flow_occ without gt_disp_change: the change of frame pair (idx-1, idx) comes from compute_gt_disp_change
    (codd.py:331-340) with the PREVIOUS frame's flow / disparity / occlusion, and the motion block uses entry [-1]."""
    from codd_b200 import ops
    monkeypatch.setattr(metrics_mod, "SequenceMetrics", _Recorder)
    seen = []

    def fake_change(flow_prev, gt_curr, gt_prev, occ_prev):
        seen.append((flow_prev, gt_curr, gt_prev, occ_prev))
        return torch.full_like(gt_prev, float(len(seen))), None

    monkeypatch.setattr(ops, "gt_disp_change", fake_change)
    model = codd_b200.build_estimator(codd_b200.codd_stereo_config(64))
    model.eval()
    B, MF, H, W = 1, 3, 64, 64
    g = torch.Generator().manual_seed(2)
    monkeypatch.setattr(model, "consistent_online_depth_estimation",
                        lambda l, r, m, s: {"pred_disp": torch.ones(B, 1, H, W), "Ts": torch.zeros(B, H, W, 7)})
    img = torch.zeros(B, MF, 3, H, W)
    gt = torch.rand(B, MF, 1, H, W, generator=g) * 60
    flow = torch.randn(B, MF, 2, H, W, generator=g)
    focc = (torch.rand(B, MF, 1, H, W, generator=g) > 0.7).float()
    metas = [[dict(img_shape=(H, W, 3), disp_range=(0.0, 64.0), intrinsics=[100.0, 100.0, 32.0, 32.0])]]
    model(return_loss=False, evaluate=True, img=[img], img_metas=metas, r_img=[img], gt_disp=[gt], gt_flow=[flow],
          gt_flow_occ=[focc])
    assert len(seen) == MF - 1
    for i, (f, gc, gp, oc) in enumerate(seen, start=1):
        assert torch.equal(f, flow[:, i - 1]) and torch.equal(gc, gt[:, i]) and torch.equal(gp, gt[:, i - 1])
        assert torch.equal(oc, focc[:, i - 1] > 0)
        c = _Recorder.calls[i]
        assert torch.equal(c["gt_disp_change"], torch.full((B, 1, H, W), float(i)))
        assert torch.equal(c["gt_flow_occ_prev"], focc[:, i - 1] > 0)
