"""Full CODD (stereo + motion + fusion) multi-frame forward through the reference-facing API
(BASELINE.json configs[2] geometry at a test size) against the composed CPU oracle:
hitnet_oracle.stereo_matching -> raft3d_oracle.motion_forward -> fusion_oracle.memory_query/update
(the call order of model/codd.py:80-126)."""
import pytest
import torch

from oracle import fusion_oracle as FO
from oracle import hitnet_oracle as O
from oracle import raft3d_oracle as R

pytestmark = pytest.mark.gpu


def gpu_argmin(stereo, left, right):
    """The CUDA path's arg-min initialisation pyramid (coarse->fine), as CPU tensors [N,h,w]."""
    from codd_b200 import ops
    with torch.no_grad():
        fl, fr = stereo.backbone.forward_pair(left.cuda(), right.cuda())
        _, hyps = stereo.tile_init(fl, fr)
    return [ops.to_nchw(h).cpu()[:, 0] for h in hyps]


def oracle_sequence(sds, lefts, rights, max_disp, intr, iters, stereo=None):
    """``stereo``: the CUDA HITNetMF module — its arg-min initialisation is handed to the oracle, which adopts the
    choices it can certify as near-ties of its own cost volume (oracle.stereo_matching_given_argmin) and fails on any
    other disagreement; one such flip would otherwise change a whole image region of every later frame."""
    hsd, msd, fsd = sds
    state, preds = {}, []
    for left, right in zip(lefts, rights):
        out = O.stereo_matching_given_argmin(hsd, left, right, max_disp, gpu_argmin(stereo, left, right))
        assert out["uncertified"] == 0, "arg-min initialisation differs beyond a near-tie"
        out = dict(pred_disp=out["pred_disp"], left_feat=out["left_feat"], right_feat=out["right_feat"], left_img=left)
        R.motion_forward(msd, "", state, out, intr, iters)
        FO.memory_query(fsd, out, state, direct=True)
        FO.memory_update(out, state)
        preds.append(out["pred_disp"])
    return torch.cat(preds, 1), out


def test_full_codd_sequence_vs_oracle():
    import codd_b200
    from codd_b200 import ops
    max_disp, iters, n, h, w, frames = 64, 2, 1, 128, 192, 3
    hsd = O.random_hitnet_params(11)
    fsd = FO.random_fusion_params(12)
    rsd = R.random_raft3d_params(13)
    model = codd_b200.build_estimator(codd_b200.codd_full_config(max_disp, iters))
    model.stereo.load_state_dict(hsd)
    model.fusion.load_state_dict(fsd)
    model.motion.raft3d.load_state_dict(rsd)
    model.cuda()
    model.eval()
    # the registry-built tree carries the reference's top-level parameter names
    keys = model.state_dict().keys()
    assert any(k.startswith("motion.raft3d.update_block.gru.convz1") for k in keys)
    assert any(k.startswith("motion.raft3d.cnet.0.stage4.1.fuse_layers.3.0.2.0") for k in keys)
    assert any(k.startswith("fusion.key_layer.2.conv1.0") for k in keys)

    # a slowly translating scene: frame t is frame 0 shifted by (t, 2t) pixels
    left0, right0 = O.synth_pair(n, h, w, max_disp, seed=21, kind="S")
    lefts = [torch.roll(left0, shifts=(t, 2 * t), dims=(2, 3)) for t in range(frames)]
    rights = [torch.roll(right0, shifts=(t, 2 * t), dims=(2, 3)) for t in range(frames)]
    intr = [float(w), float(w), w / 2.0, h / 2.0]
    metas = [[dict(min_disp=1, max_disp=max_disp, ori_shape=(h - 8, w - 2), img_shape=(h - 8, w - 2), intrinsics=intr)]]
    img, r_img = torch.stack(lefts, 1).cuda(), torch.stack(rights, 1).cuda()
    res = model(return_loss=False, rescale=True, evaluate=False, img=[img], img_metas=metas, r_img=[r_img])
    assert isinstance(res, list) and res[0].shape == (n, frames, h - 8, w - 2)

    msd = {"raft3d." + k: v for k, v in rsd.items()}
    ref, ref_out = oracle_sequence((hsd, msd, fsd), lefts, rights, max_disp, intr, iters, stereo=model.stereo)
    ref = ref[:, :, :h - 8, :w - 2]
    got = res[0].cpu()
    for t in range(frames):
        err = (got[:, t] - ref[:, t]).abs()
        frac = (err <= 1e-3 * ref[:, t].abs().clamp(min=1.0)).float().mean().item()
        print(f"frame {t}: {frac * 100:.3f}% of pixels within 1e-3, max abs err {err.max().item():.3e}")
        # frame 0 is the stereo network alone; later frames add the splat warp + fusion blend, whose
        # discrete steps (z-order, top-8 cut, warp>0 masks) can flip on fp32 rounding at isolated pixels
        assert frac >= (0.995 if t == 0 else 0.97)
    st = model.inference_state
    assert len(st["memory"]) == 3 and st["memory"][1].shape == (n, 32, h // 4, w // 4)
    assert ops.LAUNCHES[0] > 0


def test_full_codd_full_size_properties():
    """BASELINE.json configs[2] geometry (960x540 padded to 576x960, D=192) at batch 1, 2 frames, 4 RAFT3D iterations:
    properties that need no CPU oracle at this size — output shape / crop, finiteness, non-negativity, run-to-run
    determinism, the 3-tuple memory left by Fusion.memory_update and the fusion weights' range."""
    import codd_b200
    from codd_b200.synth import synth_pair
    torch.manual_seed(3)
    model = codd_b200.build_estimator(codd_b200.codd_full_config(192, iters=4)).cuda()
    model.eval()
    left, right = synth_pair(1, 576, 960, 192, seed=5, kind="S")
    img = torch.stack([left, torch.roll(left, shifts=(1, 2), dims=(2, 3))], 1).cuda()
    r_img = torch.stack([right, torch.roll(right, shifts=(1, 2), dims=(2, 3))], 1).cuda()
    metas = [[dict(min_disp=1, max_disp=192, ori_shape=(540, 960), img_shape=(540, 960),
                   intrinsics=[1050.0, 1050.0, 480.0, 270.0])]]
    run = lambda: model(return_loss=False, rescale=True, evaluate=False, img=[img], img_metas=metas, r_img=[r_img])[0]
    a = run()
    assert a.shape == (1, 2, 540, 960) and torch.isfinite(a).all() and (a >= 0).all()
    mem = model.inference_state["memory"]
    assert len(mem) == 3 and mem[1].shape == (1, 32, 144, 240) and mem[2].shape == (1, 576, 960)
    b = run()
    assert torch.equal(a, b), "full CODD forward is not deterministic"


def test_evaluate_true_through_model_vs_metrics_oracle():
    """model(..., evaluate=True, gt_*=...) on the GPU for a 3-frame full-CODD sequence (codd.py:313-355, 435-575): the
    returned meters against oracle/metrics_oracle.py (pinned against the reference's utils) evaluated on the model's own
    per-frame outputs (recorded from consistent_online_depth_estimation while the call runs)."""
    import numpy as np
    import codd_b200
    from oracle import metrics_oracle as M
    torch.manual_seed(3)
    D, h, w, H, W, T = 64, 120, 180, 128, 192, 3
    model = codd_b200.build_estimator(codd_b200.codd_full_config(D, 1)).cuda()
    model.eval()
    left, right = O.synth_pair(1, H, W, D, seed=11, kind="S")
    lefts = [torch.roll(left, shifts=(t, 2 * t), dims=(2, 3)) for t in range(T)]
    rights = [torch.roll(right, shifts=(t, 2 * t), dims=(2, 3)) for t in range(T)]
    g = torch.Generator().manual_seed(5)
    gt = torch.rand(1, T, 1, H, W, generator=g) * 70 - 3          # some invalid (<= 0) pixels
    flow = torch.randn(1, T, 2, H, W, generator=g) * 1.5
    dc = torch.randn(1, T, 1, H, W, generator=g)
    occ = (torch.rand(1, T, 1, H, W, generator=g) > 0.85).float()  # gt_disp_occ: > 0 = occluded
    intr = [float(W), float(W), W / 2.0, H / 2.0]
    rng = (0.0, float(D))
    metas = [[dict(min_disp=1, max_disp=D, ori_shape=(h, w), img_shape=(h, w), intrinsics=intr, disp_range=rng)]]
    rec = []
    inner = model.consistent_online_depth_estimation

    def recording(l, r, m, state):
        out = inner(l, r, m, state)
        rec.append(dict(pred=out["pred_disp"].detach().clone(), Ts=None if out.get("Ts") is None else out["Ts"].detach().clone()))
        return out

    model.consistent_online_depth_estimation = recording
    res = model(return_loss=False, rescale=True, evaluate=True, img=[torch.stack(lefts, 1).cuda()], img_metas=metas,
                r_img=[torch.stack(rights, 1).cuda()], gt_disp=[gt.cuda()], gt_flow=[flow.cuda()],
                gt_disp_change=[dc.cuda()], gt_disp_occ=[occ.cuda()])[0]
    assert len(rec) == T and all(r["Ts"] is not None for r in rec[1:])
    got = {k: float(v) for k, v in res.items()}

    # ---- the reference's bookkeeping on the CPU, frame by frame
    exp = {k: [] for k in ("epe", "th3", "tepe", "tepe_rel", "th1_tepe_rel", "th3_tepe", "flow_mag")}
    sf = np.zeros(5)
    prev = None
    intr_np = np.array([intr], np.float32)
    for t in range(T):
        pred = rec[t]["pred"][:, :, :h, :w].cpu().numpy()
        gtt = gt[:, t, :, :h, :w].numpy()
        seg = (occ[:, t] <= 0)[:, :, :h, :w].float().numpy()
        mask = M.valid_mask(gtt, rng, seg=seg)
        do = M.disp_metrics(pred, gtt, mask)
        if do["n"]:
            exp["epe"].append(do["epe"]); exp["th3"].append(do["th3"])
        if prev is not None:
            to = M.temporal_metrics(prev["flow"], gtt, pred, seg, prev["gt"], prev["pred"], prev["mask"], rng)
            exp["flow_mag"].append(to["flow_mag"])
            if to["updated"]:
                for k in ("tepe", "tepe_rel", "th1_tepe_rel", "th3_tepe"):
                    exp[k].append(to[k])
            # provided gt_disp_change: the reference uses entry [-2], i.e. the previous frame's (codd.py:519-540)
            Ts = rec[t]["Ts"][:, :h, :w].cpu().numpy()
            o = M.sceneflow_metrics(Ts, prev["pred"], intr_np, prev["flow"], dc[:, t - 1, :, :h, :w].numpy(), prev["gt"], rng,
                                    seg=seg)
            sf += np.array([o["n"], o["sum_sf"], o["sum_of"], o["n1_sf"], o["n1_of"]], np.float64)
        prev = dict(flow=flow[:, t, :, :h, :w].numpy(), gt=gtt, pred=pred, mask=mask)
    for k, v in exp.items():
        assert v, k
        assert got[k] == pytest.approx(float(np.mean(v)), rel=1e-6), k      # the meters come back as float32 tensors
    assert got["count"] == sf[0] and sf[0] > 0
    assert got["epe2d_scene_flow"] == pytest.approx(sf[1], rel=1e-4)
    assert got["epe2d_optical_flow"] == pytest.approx(sf[2], rel=1e-4)
    assert abs(got["1px_scene_flow"] - sf[3]) <= 3 and abs(got["1px_optical_flow"] - sf[4]) <= 3
