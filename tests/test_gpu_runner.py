"""N2: the sequence driver (CUDA-graph replay per shape, alternating serving slots) against the CPU oracle and the eager
per-frame path; reference npz layout; per-sequence evaluation statistics."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def make_sequences(seed, n_seq, t, h, w):
    g = np.random.default_rng(seed)
    seqs = []
    for s in range(n_seq):
        left = g.integers(0, 256, (t, h, w, 3), dtype=np.uint8)
        right = np.roll(left, -3, axis=2)
        gt = g.uniform(1, 40, (t, 1, h, w)).astype(np.float32)
        flow = g.normal(0, 1.5, (t, 2, h, w)).astype(np.float32)
        seqs.append(dict(name=f"seq{s}", left=left, right=right, gt_disp=gt, gt_flow=flow))
    return seqs


def _oracle_disparity(model, sd, l8, r8, max_disp):
    """The CPU oracle's answer for one uint8 frame pair: staging oracle (N1) -> HITNetMF oracle, adopting the CUDA
    path's arg-min only where the oracle's own cost volume certifies a near-tie."""
    from codd_b200 import ops
    from oracle import hitnet_oracle as O
    from oracle import staging_oracle as S
    left, right = torch.from_numpy(S.stage_images_u8(l8)), torch.from_numpy(S.stage_images_u8(r8))
    with torch.no_grad():
        fl, fr = model.stereo.backbone.forward_pair(left.cuda(), right.cuda())
        _, hyps = model.stereo.tile_init(fl, fr)
    ref = O.stereo_matching_given_argmin(sd, left, right, max_disp, [ops.to_nchw(h).cpu()[:, 0] for h in hyps])
    assert ref["uncertified"] == 0
    return ref["pred_disp"]


@pytest.mark.parametrize("use_graph", [True, False])
def test_runner_matches_oracle(tmp_path, use_graph):
    """Graph-replayed sequence driver vs the CPU ORACLE (staging + stereo), per frame; the eager CUDA path is also
    required to be bit-identical (a replay bug would show there first); one npz per sequence in the reference's layout;
    metrics per sequence (meters reset at every sequence start) against a fresh SequenceMetrics fed with the oracle-
    checked disparities."""
    import codd_b200
    from codd_b200 import ops
    from codd_b200.metrics import SequenceMetrics
    from codd_b200.runner import StereoSequenceRunner
    torch.manual_seed(0)
    model = codd_b200.build_estimator(codd_b200.codd_stereo_config(64)).cuda()
    model.eval()      # (the reference's train() override returns None, so no chaining)
    sd = {k[len("stereo."):]: v.detach().cpu() for k, v in model.state_dict().items()}
    seqs = make_sequences(5, n_seq=3, t=5, h=70, w=100)
    seqs[2] = make_sequences(6, 1, 3, 96, 130)[0] | dict(name="seq2")     # a second input shape
    for s in seqs:
        s["disp_range"] = (0.0, 64.0)
    runner = StereoSequenceRunner(model, use_graph=use_graph, n_streams=2)
    res = runner.run_sequences(seqs, out_dir=str(tmp_path), metrics=SequenceMetrics((0.0, 64.0), max_frames=2), batch=2)
    stats = res.pop("__stats__")
    assert set(res) == {"seq0", "seq1", "seq2"}
    # input shapes seen: (2|1, 70, 100) and (2|1, 96, 130), two serving slots each
    assert runner.graphs_captured == (8 if use_graph else 0)
    worst = 1.0
    with torch.no_grad():
        for seq in seqs:
            T = seq["left"].shape[0]
            f = os.path.join(str(tmp_path), seq["name"] + ".disp.pred.npz")
            saved = np.load(f)["disp"]
            assert saved.shape == (1, T) + seq["left"].shape[1:3]          # [B=1, MF, H, W] as show_result writes it
            assert np.array_equal(saved[0], res[seq["name"]][:, 0])
            for t in range(T):
                l8, r8 = seq["left"][t:t + 1], seq["right"][t:t + 1]
                h, w = l8.shape[1:3]
                eager = model.stereo.stereo_matching(ops.stage_images_u8(torch.from_numpy(l8).cuda()),
                                                     ops.stage_images_u8(torch.from_numpy(r8).cuda()))["pred_disp"]
                assert np.array_equal(res[seq["name"]][t:t + 1], eager[:, :, :h, :w].cpu().numpy()), (seq["name"], t)
                if t in (0, T - 1):                                        # oracle: first and last frame of a sequence
                    ref = _oracle_disparity(model, sd, l8, r8, 64)[:, :, :h, :w]
                    got = torch.from_numpy(res[seq["name"]][t:t + 1])
                    ok = ((got - ref).abs() <= 1e-3 * ref.abs().clamp(min=1.0)).float().mean().item()
                    worst = min(worst, ok)
                    assert ok >= 0.995, (seq["name"], t, ok)
    print(f"runner vs oracle: worst frame has {worst*100:.3f}% of pixels within 1e-3")
    # per-sequence evaluation: one row per sequence, each equal to a fresh meter over that sequence alone
    assert stats.n == 3 and stats.names == ["seq0", "seq1", "seq2"]
    for seq, row in zip(seqs, stats.rows):
        sm = SequenceMetrics((0.0, 64.0), max_frames=seq["left"].shape[0])
        for t in range(seq["left"].shape[0]):
            sm.update(torch.from_numpy(res[seq["name"]][t:t + 1]).cuda(), torch.from_numpy(seq["gt_disp"][t:t + 1]).cuda(),
                      gt_flow=torch.from_numpy(seq["gt_flow"][t:t + 1]).cuda())
        alone = sm.collect()
        assert row == alone and np.isfinite(row["epe"]) and row["epe"] > 0 and np.isfinite(row["flow_mag"])
    assert abs(stats.mean["epe"] - np.mean([r["epe"] for r in stats.rows])) < 1e-12
    assert abs(stats.std["epe"] - np.std([r["epe"] for r in stats.rows])) < 1e-12
    stats.dump(os.path.join(str(tmp_path), "stats.csv"))
    assert open(os.path.join(str(tmp_path), "stats.csv")).read().count("\n") == 6


def test_runner_sharding_covers_all_sequences():
    import codd_b200
    from codd_b200.runner import StereoSequenceRunner
    torch.manual_seed(0)
    model = codd_b200.build_estimator(codd_b200.codd_stereo_config(64)).cuda()
    model.eval()      # (the reference's train() override returns None, so no chaining)
    seqs = make_sequences(7, n_seq=3, t=2, h=64, w=64)
    runner = StereoSequenceRunner(model, n_streams=2)
    r0 = runner.run_sequences(seqs, rank=0, world_size=2)
    r1 = runner.run_sequences(seqs, rank=1, world_size=2)
    assert set(r0) == {"seq0", "seq2"} and set(r1) == {"seq1"}
    full = runner.run_sequences(seqs)
    assert "__stats__" not in full      # no ground truth consumer asked for -> no statistics object
    for k in full:
        assert np.array_equal(full[k], (r0 | r1)[k])
