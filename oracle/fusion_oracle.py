"""CPU oracle for the Fusion forward.

TEST INFRASTRUCTURE ONLY (see oracle/hitnet_oracle.py for the rules).  Functional torch-CPU
restatement of model/fusion/fusion.py:357-410 (memory_query / memory_update) and of the helpers
it calls (utils/warp.py:43-66 disp_warp; fusion.py:168-355), over a flat state_dict with the
reference's parameter names.  Pinned against the unmodified reference module in
tests/test_oracle_vs_reference.py (tolerance 1e-6: the reference reduces the channel dot-products
through an unfold + sum whose vectorisation order is not restated; everything else is exact).
"""
import math

import torch
import torch.nn.functional as F

from .hitnet_oracle import l1_over_channels, warp_right_direct, warp_right


def _conv(sd, name, x, padding=0):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), padding=padding)


def key_layer(sd, left_feat):
    """fusion.py:74-80: 1x1 -> ReLU -> BasicBlock(3x3 Mish, 3x3, +x) -> ReLU -> 1x1."""
    x = F.relu(_conv(sd, "key_layer.0", left_feat))
    y = F.mish(_conv(sd, "key_layer.2.conv1.0", x, padding=1))
    y = _conv(sd, "key_layer.2.conv2", y, padding=1)
    x = F.relu(y + x)
    return _conv(sd, "key_layer.4", x)


def patch_shifts(x, dilation=2):
    """The 9 zero-padded shifts of a 3x3 patch with dilation 2, in unfold order (row-major)."""
    n, c, h, w = x.shape
    p = F.pad(x, (dilation, dilation, dilation, dilation))
    out = []
    for py in range(3):
        for px in range(3):
            out.append(p[:, :, py * dilation:py * dilation + h, px * dilation:px * dilation + w])
    return out


def px2patch_corr(q, mem, self_corr=False):
    """fusion.py:168-198.  C == 1: pixel minus patch; else channel dot-product; / sqrt(C)."""
    c = q.shape[1]
    vals = []
    for idx, s in enumerate(patch_shifts(mem)):
        if self_corr and idx == 4:
            continue
        vals.append((q - s) if c == 1 else (q * s).sum(1, keepdim=True))
    return torch.cat(vals, 1) / math.sqrt(c)


def disparity_confidence(pred_curr, pred_warp, fea_l, fea_r, ds=4, in_channels=24, direct=False):
    """fusion.py:200-241: +-1 local stereo costs of the current and the warped disparity at 1/4
    resolution (disp_warp = the same normalise / grid_sample round trip as HITNet's warp)."""
    o = ds // 2 - 1
    pc, pw = pred_curr[..., o::ds, o::ds], pred_warp[..., o::ds, o::ds]
    wfn = warp_right_direct if direct else warp_right
    cv_pred, cv_warp = [], []
    for k in (-1, 0, 1):
        cv_warp.append(l1_over_channels(fea_l - wfn(fea_r, pw / ds + k)) / (in_channels / 24.0))
        cv_pred.append(l1_over_channels(fea_l - wfn(fea_r, pc / ds + k)) / (in_channels / 24.0))
    return torch.cat(cv_pred, 1), torch.cat(cv_warp, 1)


def compute_input_cues(pred_curr, pred_warp, feat_curr, feat_warp, flow_warp, conf_warp, fea_l, fea_r, direct=False):
    """fusion.py:243-318 -> (corr_feat [N,31,h,w], corr_feat_fr [N,32,H,W])."""
    cost_curr, cost_warp = disparity_confidence(pred_curr, pred_warp, fea_l, fea_r, direct=direct)
    feat_cross = px2patch_corr(feat_curr, feat_warp)
    feat_self = torch.cat([px2patch_corr(feat_curr, feat_curr, True), px2patch_corr(feat_warp, feat_warp, True)], 1)
    disp_cross = px2patch_corr(pred_curr, pred_warp).abs()
    disp_self = torch.cat([px2patch_corr(pred_curr, pred_curr, True), px2patch_corr(pred_warp, pred_warp, True)], 1).abs()
    corr_feat = torch.cat([feat_cross, feat_self, cost_curr, cost_warp], 1)
    corr_feat_fr = torch.cat([disp_cross, disp_self, flow_warp, (pred_warp > 0).float(), conf_warp], 1)
    return corr_feat, corr_feat_fr


def fuse(sd, corr_feat, pred_curr, pred_warp, feat_curr, ds=4):
    """fusion.py:320-355 -> fusion weights at full resolution."""
    o = ds // 2 - 1
    pc, pw = pred_curr[..., o::ds, o::ds], pred_warp[..., o::ds, o::ds]
    corr = F.relu(_conv(sd, "conv_corr.2", F.relu(_conv(sd, "conv_corr.0", corr_feat))))
    disp = F.relu(_conv(sd, "conv_disp.2", F.relu(_conv(sd, "conv_disp.0", torch.cat([pc, pw], 1), 3)), 1))
    mo = F.relu(_conv(sd, "motion_conv.0", torch.cat([corr, disp], 1), 3))
    net = F.relu(_conv(sd, "residual_conv.0", torch.cat([feat_curr, mo, pc, pw], 1), 1)) + corr
    w = torch.sigmoid(_conv(sd, "weight_head.1", _conv(sd, "weight_head.0", net, 1)))
    return F.interpolate(w, scale_factor=ds)


def forget_head(sd, corr_feat_fr):
    x = _conv(sd, "forget_head.2", _conv(sd, "forget_head.1", _conv(sd, "forget_head.0", corr_feat_fr), 1))
    return torch.sigmoid(x)


def memory_query(sd, outputs, state, direct=False):
    """fusion.py:357-402.  Mutates ``outputs`` exactly like the reference."""
    left_feat, pred_curr = outputs["left_feat"], outputs["pred_disp"]
    feat_curr = key_layer(sd, left_feat)
    if "memory" not in state:
        outputs["left_feat"] = feat_curr
        return outputs
    _, feat_warp, conf_warp, pred_warp, flow_warp = state["memory"]
    corr_feat, corr_feat_fr = compute_input_cues(pred_curr, pred_warp, feat_curr, feat_warp, flow_warp, conf_warp,
                                                 outputs["left_feat"], outputs["right_feat"], direct)
    mask = (pred_warp > 0.0).float()
    wf = fuse(sd, corr_feat, pred_curr, pred_warp, feat_curr) * mask
    wr = forget_head(sd, corr_feat_fr) * mask
    outputs["pred_disp"] = pred_curr * (1 - wf * wr) + pred_warp * wf * wr
    outputs["fusion_weights"], outputs["reset_weights"] = wf, wr
    outputs["pred_curr"], outputs["pred_warp"] = pred_curr, pred_warp
    outputs["left_feat"] = feat_curr
    outputs["_corr_feat"], outputs["_corr_feat_fr"] = corr_feat, corr_feat_fr
    return outputs


def memory_update(outputs, state):
    """fusion.py:404-410."""
    state["memory"] = [outputs["left_img"], outputs["left_feat"], outputs["pred_disp"].squeeze(1)]


def fusion_param_shapes():
    """(name, cout, cin, k) — SURVEY.md Appendix B, fusion block (in_channels=24, fusion_channel=32)."""
    return [("key_layer.0", 32, 24, 1), ("key_layer.2.conv1.0", 32, 32, 3), ("key_layer.2.conv2", 32, 32, 3),
            ("key_layer.4", 32, 32, 1), ("conv_corr.0", 64, 31, 1), ("conv_corr.2", 32, 64, 1),
            ("conv_disp.0", 32, 2, 7), ("conv_disp.2", 32, 32, 3), ("motion_conv.0", 30, 64, 7),
            ("weight_head.0", 32, 32, 3), ("weight_head.1", 1, 32, 1), ("forget_head.0", 16, 32, 1),
            ("forget_head.1", 8, 16, 3), ("forget_head.2", 1, 8, 1), ("residual_conv.0", 32, 64, 3)]


def random_fusion_params(seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, co, ci, k in fusion_param_shapes():
        bound = 1.0 / math.sqrt(ci * k * k)
        sd[name + ".weight"] = (torch.rand(co, ci, k, k, generator=g) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand(co, generator=g) * 2 - 1) * bound
    return sd


def synth_fusion_inputs(n, h, w, seed=0, max_disp=64.0):
    """Synthetic (outputs, state) for a t>=1 frame: H, W multiples of 4."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    pred = (torch.rand(n, 1, h, w, generator=g) * max_disp)
    pred_warp = (pred + r(n, 1, h, w) * 2).clamp(min=0)
    pred_warp[torch.rand(n, 1, h, w, generator=g) < 0.1] = 0.0        # holes of the splat warp
    outputs = dict(pred_disp=pred, left_feat=r(n, 24, h // 4, w // 4), right_feat=r(n, 24, h // 4, w // 4),
                   left_img=r(n, 3, h, w))
    state = dict(memory=[r(n, 3, h, w), r(n, 32, h // 4, w // 4), torch.rand(n, 3, h, w, generator=g), pred_warp,
                         r(n, 3, h, w)])
    return outputs, state
