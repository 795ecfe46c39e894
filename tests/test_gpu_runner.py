"""N2: the sequence driver (CUDA-graph replay per shape, alternating serving slots) returns exactly what the eager
per-frame path returns, writes the reference's npz files and feeds the on-GPU meters."""
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


@pytest.mark.parametrize("use_graph", [True, False])
def test_runner_matches_eager(tmp_path, use_graph):
    import codd_b200
    from codd_b200 import ops
    from codd_b200.metrics import SequenceMetrics
    from codd_b200.runner import StereoSequenceRunner
    torch.manual_seed(0)
    model = codd_b200.build_estimator(codd_b200.codd_stereo_config(64)).cuda()
    model.eval()      # (the reference's train() override returns None, so no chaining)
    seqs = make_sequences(5, n_seq=3, t=5, h=70, w=100)
    seqs[2] = make_sequences(6, 1, 3, 96, 130)[0] | dict(name="seq2")     # a second input shape
    runner = StereoSequenceRunner(model, use_graph=use_graph, n_streams=2)
    sm = SequenceMetrics((0.0, 64.0), max_frames=64)
    res = runner.run_sequences(seqs, out_dir=str(tmp_path), metrics=sm, batch=2)
    assert set(res) == {"seq0", "seq1", "seq2"}
    # input shapes seen: (2|1, 70, 100) and (2|1, 96, 130), two serving slots each
    assert runner.graphs_captured == (8 if use_graph else 0)
    with torch.no_grad():
        for seq in seqs:
            for t in range(seq["left"].shape[0]):
                l8 = torch.from_numpy(seq["left"][t:t + 1]).cuda()
                r8 = torch.from_numpy(seq["right"][t:t + 1]).cuda()
                h, w = l8.shape[1:3]
                ref = model.stereo.stereo_matching(ops.stage_images_u8(l8), ops.stage_images_u8(r8))["pred_disp"]
                ref = ref[:, :, :h, :w].cpu().numpy()
                assert np.array_equal(res[seq["name"]][t:t + 1], ref), (seq["name"], t)
                f = os.path.join(str(tmp_path), seq["name"], f"{t:06d}.disp.pred.npz")
                assert np.array_equal(np.load(f)["disp"], ref)
    out = sm.collect()
    assert sm.frames == 13 and np.isfinite(out["epe"]) and out["epe"] > 0 and np.isfinite(out["flow_mag"])


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
    for k in full:
        assert np.array_equal(full[k], (r0 | r1)[k])
