"""Fusion — drop-in for model/fusion/fusion.py:41-449 (registry name ``Fusion``), inference path.

Same constructor (``in_channels, fusion_channel, loss, corr_cfg, ds_scale``), parameter tree
(key_layer, conv_corr, conv_disp, motion_conv, weight_head, forget_head, residual_conv) and the
``memory_query(outputs, state)`` / ``memory_update(outputs, state)`` contract, including the
3-tuple / 5-tuple ``state["memory"]`` convention (fusion.py:406-410 vs motion.py:207).
"""
import torch
import torch.nn as nn

from .. import ops
from ..lib import ACT_MISH, ACT_NONE, ACT_RELU, ACT_SIGMOID
from ..registry import MODELS, build_loss
from ..stereo._params import PackedWeights, run_conv


class GradientClip(nn.Module):
    """Identity in the forward pass (model/motion/raft3d/raft3d.py GradientClip)."""

    def forward(self, x):
        return x


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, c1, c2, s, p, d):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv2d(c1, c2, kernel_size=3, stride=s, padding=d if d > 1 else p, dilation=d),
                                   nn.Mish(inplace=True))
        self.conv2 = nn.Conv2d(c2, c2, kernel_size=3, stride=1, padding=d if d > 1 else p, dilation=d)


@MODELS.register_module(force=True)
class Fusion(nn.Module):
    def __init__(self, in_channels, fusion_channel, loss=None, corr_cfg=dict(), ds_scale=4):
        super().__init__()
        self.loss = build_loss(loss) if loss is not None else None
        self.fusion_channel = fusion_channel
        self.ds_scale = ds_scale
        self.in_channels = in_channels
        self.patch_size = corr_cfg.get("patch_size", 3)
        if self.patch_size != 3 or fusion_channel != 32:
            raise NotImplementedError("codd_b200.Fusion implements the reference configuration (patch 3, 32 channels)")
        fc = fusion_channel
        self.key_layer = nn.Sequential(
            nn.Conv2d(in_channels, fc, 1, 1, 0, 1), nn.ReLU(inplace=True), BasicBlock(fc, fc, s=1, p=1, d=1),
            nn.ReLU(inplace=True), nn.Conv2d(fc, fc, 1, 1, 0, 1))
        self.conv_corr = nn.Sequential(nn.Conv2d(16 + 9 + 6, fc * 2, 1, padding=0, bias=True), nn.ReLU(inplace=True),
                                       nn.Conv2d(fc * 2, fc, 1, padding=0, bias=True), nn.ReLU(inplace=True))
        self.conv_disp = nn.Sequential(nn.Conv2d(2, fc, 7, padding=3), nn.ReLU(inplace=True),
                                       nn.Conv2d(fc, fc, 3, padding=1, bias=True), nn.ReLU(inplace=True))
        self.motion_conv = nn.Sequential(nn.Conv2d(fc * 2, fc - 2, 7, padding=3, bias=True), nn.ReLU(inplace=True))
        self.weight_head = nn.Sequential(nn.Conv2d(fc, fc, 3, padding=1, bias=True), nn.Conv2d(fc, 1, 1, padding=0, bias=True),
                                         GradientClip(), nn.Sigmoid())
        self.forget_head = nn.Sequential(nn.Conv2d(6 + 16 + 9 + 1, 16, 1, padding=0, bias=True),
                                         nn.Conv2d(16, 8, 3, padding=1, bias=True), nn.Conv2d(8, 1, 1, padding=0, bias=True),
                                         GradientClip(), nn.Sigmoid())
        self.residual_conv = nn.Sequential(nn.Conv2d(fc + fc, fc, 3, padding=1, bias=True), nn.ReLU(inplace=True))
        self._pw = PackedWeights()
        n_parameters = sum(p.numel() for n, p in self.named_parameters())
        print("PARAM STATUS: total number of parameters %.3fM in fusion network" % (n_parameters / 1000 ** 2))

    # -- pieces ------------------------------------------------------------------------------
    def _c(self, conv, x, act, **kw):
        return run_conv(self._pw, conv, x, act, **kw)

    def _key(self, left_feat, out=None):
        k = self.key_layer
        x = self._c(k[0], ops.to_nhwc(left_feat), ACT_RELU)
        y = self._c(k[2].conv1[0], x, ACT_MISH)
        x = self._c(k[2].conv2, y, ACT_RELU, residual=x)       # BasicBlock: conv2(..) + x, then the ReLU after it
        wp, b = self._pw.conv(k[4])
        return ops.conv2d(x, wp, b, self.fusion_channel, 1, act=ACT_NONE, out=out)

    def memory_query(self, outputs, state, *args, **kwargs):
        left_feat, pred_curr = outputs["left_feat"], outputs["pred_disp"]
        if "memory" not in state:
            outputs["left_feat"] = self._key(left_feat)
            return
        left_img_prev, feat_warp, confidence_warp, pred_warp, flow_warp = state["memory"]
        if pred_warp.dim() == 3:
            pred_warp = pred_warp.unsqueeze(1)
        n, _, H, W = pred_curr.shape
        ds = self.ds_scale
        dev = pred_curr.device
        # [feat_curr(32) | motion(30) | pred_curr, pred_warp (2)]: residual_conv's input, written in place
        inp = ops.empty_nhwc(n, 64, H // ds, W // ds, dev)
        feat_curr = self._key(left_feat, out=inp[:, :32])
        corr_feat, disp2 = ops.fusion_cues_lowres(ops.to_nhwc(feat_curr), ops.to_nhwc(feat_warp), ops.to_nhwc(left_feat),
                                                  outputs["right_feat"], pred_curr, pred_warp, ds, extra=inp[:, 62:64])
        # fuse (fusion.py:320-355)
        corr = self._c(self.conv_corr[2], self._c(self.conv_corr[0], corr_feat, ACT_RELU), ACT_RELU)
        disp = self._c(self.conv_disp[2], self._c(self.conv_disp[0], disp2, ACT_RELU), ACT_RELU)
        wp, b = self._pw.conv(self.motion_conv[0])
        ops.conv2d(corr, wp, b, 30, 7, 1, 3, 1, ACT_RELU, x2=disp, out=inp[:, 32:62])
        wp, b = self._pw.conv(self.residual_conv[0])
        net = ops.conv2d(inp, wp, b, 32, 3, 1, 1, 1, ACT_RELU, residual=corr, res_after_act=True)   # relu(conv) + corr
        wf_lr = self._c(self.weight_head[1], self._c(self.weight_head[0], net, ACT_NONE), ACT_SIGMOID)
        # forget head on the full-resolution cues (fusion.py:123-132, 387)
        w0, b0 = self._pw.raw(self.forget_head[0])
        r16 = ops.fusion_forget_in(pred_curr, pred_warp, flow_warp, confidence_warp, w0, b0)
        r8 = self._c(self.forget_head[1], r16, ACT_NONE)
        w2, b2 = self._pw.raw(self.forget_head[2])
        fused, wf, wr = ops.fusion_blend(pred_curr, pred_warp, r8, w2, b2, wf_lr, ds)
        outputs["pred_disp"] = fused
        outputs["fusion_weights"] = wf
        outputs["reset_weights"] = wr
        outputs["pred_curr"] = pred_curr
        outputs["pred_warp"] = pred_warp
        outputs["left_feat"] = feat_curr

    def memory_update(self, outputs, state, *args, **kwargs):
        state["memory"] = [outputs["left_img"], outputs["left_feat"], outputs["pred_disp"].squeeze(1)]

    def losses(self, *args, **kwargs):
        raise NotImplementedError("codd_b200 is a forward-only build (training losses: out of scope, DESIGN.md)")

    def freeze(self):
        self.eval()
        if self.loss is not None:
            self.loss.eval()
        for param in self.parameters():
            param.requires_grad = False
