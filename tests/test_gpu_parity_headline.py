"""Oracle parity AT THE BASELINE CONFIG: 576x960 D=192 (BASELINE.json configs[1] geometry, one pair)
and 384x1280 D=192 (configs[3] geometry).  The CPU oracle (pinned bitwise against the unmodified reference,
tests/test_oracle_vs_reference.py) runs once per size (a few seconds on the box's host cores) and the CUDA path is
compared stage by stage:

  * backbone features of all five levels (model/stereo/hitnet/backbone.py:69-88)           max abs / rel error
  * the five arg-min maps (initialization.py:158-225)                                      bit-exact count; every
    disagreement must be EXPLAINED by the measured rounding of the tile features: with e = max |tile feature error| of
    that level, a cost moves by at most 16 * 2e, so two costs can swap order only if the oracle's gap is <= 64 e
  * the arg-max hypothesis select of tile_update1..4 (propagation.py:225-248)              flips per level, given
    the same arg-min choices on both sides
  * pred_disp (hitnet.py:75-100)                                                           1e-3 * max(1,|d|)

The measured statistics are written to ``gpurun_out/parity_r02_<tag>.json`` (committed as profiles/parity_r02.json);
the assertions below are those measurements plus a small margin, not a blanket allowance."""
import json
import os

import pytest
import torch

from oracle import hitnet_oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# measured on B200 at HEAD (profiles/parity_r02.json); margins are ~2x the measurement
BOUNDS = {
    #            argmin flips/level   select flips/level, |conf gap| certified   min fraction within 1e-3, max abs error
    # measured r02 (profiles/parity_r02.json, final state): 540p 0/0/2/2/0 arg-min flips per level among 135..34560 tiles
    # (every one a certified near-tie: gap below what the measured tile-feature rounding explains), kitti 0/0/1/1/0; no select flip at either size; 100 % of pixels within 1e-3 given those choices, max abs 1.2e-4
    "540p": dict(argmin_flips=4, select_flips=2, conf=1e-5, frac=0.9999, max_abs=1e-3),
    "kitti": dict(argmin_flips=4, select_flips=2, conf=1e-5, frac=0.9999, max_abs=1e-3),
}


def _oracle_levels(sd, fl, fr, hyps, other_select=None, conf_tol=0.0):
    """Reference propagation with the per-level select decisions exposed.  ``other_select`` (per level, bool [N,h,w],
    True = current hypothesis taken): where it disagrees with the oracle's arg-max and the oracle's two confidences lie
    within ``conf_tol`` of each other, the other implementation's choice is adopted (certified near-tie) so that the
    levels below are compared on identical discrete decisions."""
    psd = O._sub(sd, "tile_update.")
    prev = O.tile_update0(O._sub(psd, "tile_update0."), fl[0], fr[0], hyps[0], True)
    stats = []
    for k in range(1, 5):
        prev, aux = O.tile_update(O._sub(psd, f"tile_update{k}."), fl[k], fr[k], hyps[k], prev, True, return_aux=True)
        if other_select is not None:
            ref_sel = aux["select"][:, 0] > 0.5
            diff = ref_sel != other_select[k]
            margin = (aux["update"][:, 0] - aux["update"][:, 1]).abs()
            near = diff & (margin <= conf_tol)
            stats.append(dict(level=k, tiles=ref_sel.numel(), flips=int(diff.sum()), uncertified=int((diff & ~near).sum()),
                              max_conf_margin=margin[diff].max().item() if diff.any() else 0.0))
            if near.any():
                sel = torch.where(near, other_select[k], ref_sel).unsqueeze(1).float()
                prev = sel * aux["cur"] + (1 - sel) * aux["prev"]
    r1 = O.post_tile_update(O._sub(psd, "tile_update4_1."), fl[2], prev, 4)
    r05 = O.post_tile_update(O._sub(psd, "tile_update5."), fl[3], O.plane_upsample(r1, 1, 2), 4)
    r025 = O.post_tile_update(O._sub(psd, "tile_update6."), fl[4], O.plane_upsample(r05, 1, 2), 2, final=True)
    return r025[:, 0:1], stats


def _within(a, b, tol=1e-3):
    err = (a - b).abs()
    return (err <= tol * b.abs().clamp(min=1.0)).float().mean().item(), err.max().item()


@pytest.mark.parametrize("tag,H,W", [("540p", 576, 960), ("kitti", 384, 1280)])
def test_headline_parity(tag, H, W):
    import codd_b200
    from codd_b200 import ops
    D = 192
    sd = O.random_hitnet_params(0)
    left, right = O.synth_pair(1, H, W, D, seed=1234, kind="S")
    torch.set_num_threads(os.cpu_count() or 1)

    # ---------------- CUDA path, stage by stage
    m = codd_b200.MODELS.build(codd_b200.hitnet_config(D))
    m.load_state_dict(sd, strict=True)
    m.cuda().eval()
    tu = m.tile_update
    for k in range(1, 5):
        getattr(tu, f"tile_update{k}").keep_aux = True
    with torch.no_grad():
        gfl, gfr = m.backbone.forward_pair(left.cuda(), right.cuda())
        _, ghyps = m.tile_init(gfl, gfr)
        gtiles = m.tile_init.tile_features(gfl, gfr)
        out = m.stereo_matching(left.cuda(), right.cuda())
        torch.cuda.synchronize()
    g_sel = [None] + [(getattr(tu, f"tile_update{k}").aux["update"][:, 1] >
                       getattr(tu, f"tile_update{k}").aux["update"][:, 0]).cpu() for k in range(1, 5)]
    g_argmin = [ops.to_nchw(h).cpu()[:, 0] for h in ghyps]
    pred = out["pred_disp"].cpu()

    # ---------------- oracle
    with torch.no_grad():
        bsd = O._sub(sd, "backbone.")
        fl, fr = O.backbone(bsd, left), O.backbone(bsd, right)
        isd = O._sub(sd, "tile_init.")
        tiles = O.tile_features(isd, fl, fr)
        hyps, cvs = O.tile_hypotheses(isd, tiles, fl, D, return_cv=True)
        ref_pred, _ = _oracle_levels(sd, fl, fr, hyps)

    stats = dict(tag=tag, shape=[1, 3, H, W], max_disp=D, features=[], argmin=[], select=[])
    # features
    for k in range(5):
        for side, g, r in (("left", gfl[k], fl[k]), ("right", gfr[k], fr[k])):
            gk = ops.to_nchw(g).cpu()
            err = (gk - r).abs()
            stats["features"].append(dict(level=k, side=side, max_abs=err.max().item(),
                                          max_rel_to_scale=(err.max() / r.abs().max()).item()))
            torch.testing.assert_close(gk, r, rtol=1e-4, atol=1e-4)
    assert torch.equal(ops.to_nchw(out["left_feat"]).cpu(), ops.to_nchw(gfl[2]).cpu())

    # arg-min maps: count flips, certify each against the oracle's cost volume, record the worst relative gap
    adopted = [h.clone() for h in hyps]
    for k in range(5):
        ref = hyps[k][:, 0]
        got = g_argmin[k]
        diff = got != ref
        n_flip = int(diff.sum())
        # measured rounding of this level's tile features (the inputs of the cost volume) and the bound it implies
        e_t = max((ops.to_nchw(gtiles[k][s]).cpu() - tiles[k][s]).abs().max().item() for s in (0, 1))
        bound = 64.0 * e_t
        gap = rel_gap = 0.0
        if n_flip:
            cmin = cvs[k].min(1)[0]
            cother = cvs[k].gather(1, got.long().clamp(0, cvs[k].shape[1] - 1).unsqueeze(1)).squeeze(1)
            gap = (cother - cmin)[diff].max().item()
            rel_gap = ((cother - cmin) / cmin.abs().clamp(min=1e-6))[diff].max().item()
            adopted[k][:, 0] = torch.where(diff, got, ref)
        stats["argmin"].append(dict(level=k, tiles=ref.numel(), flips=n_flip, max_abs_gap=gap, max_rel_gap=rel_gap,
                                    tile_feature_max_abs_err=e_t, explained_up_to=bound))

    # propagation on the same arg-min choices: select flips per level (certified by the oracle's confidence margin and
    # adopted, so every level is compared on identical upstream decisions), then the final disparity
    with torch.no_grad():
        ref_pred2, stats["select"] = _oracle_levels(sd, fl, fr, adopted, other_select=g_sel, conf_tol=BOUNDS[tag]["conf"])

    frac_raw, mx_raw = _within(pred, ref_pred)
    frac, mx = _within(pred, ref_pred2)
    stats["pred_disp"] = dict(frac_within_1e3_vs_oracle=frac_raw, max_abs_vs_oracle=mx_raw,
                              frac_within_1e3_given_argmin=frac, max_abs_given_argmin=mx,
                              argmin_flips_total=sum(a["flips"] for a in stats["argmin"]),
                              select_flips_total=sum(s["flips"] for s in stats["select"]))
    print(json.dumps(stats))
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, f"parity_r02_{tag}.json"), "w") as f:
        json.dump(stats, f, indent=1)
    B = BOUNDS[tag]
    for a in stats["argmin"]:
        assert a["flips"] <= B["argmin_flips"], f"level {a['level']}: {a['flips']} arg-min flips"
        assert a["max_abs_gap"] <= a["explained_up_to"], \
            f"level {a['level']}: arg-min disagreement (gap {a['max_abs_gap']:.3e}) exceeds what the feature rounding explains"
    for a in stats["select"]:
        assert a["flips"] <= B["select_flips"] and a["uncertified"] == 0, f"hypothesis select, level {a['level']}: {a}"
    assert frac >= B["frac"] and mx <= B["max_abs"], f"pred_disp: only {frac*100:.3f}% within 1e-3 (max abs {mx:.3e})"
