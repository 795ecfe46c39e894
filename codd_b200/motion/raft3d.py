"""RAFT3D — drop-in for model/motion/raft3d/raft3d.py:43-280 (registry name ``RAFT3D``), inference path.

Same constructor (``cnet_cfg``), parameter tree (fnet.*, cnet.0.* HRNet, cnet.1.convs.0, update_block.{gru,
corr_enc,flow_enc,ae,delta,weight,mask}) and ``forward(image_curr, depth_prev, depth_curr, intrinsics, state,
outputs, iters, train_mode)`` contract: first frame stores ``state["raft_feat"/"raft_netinp"]``, later frames
write ``outputs["Ts"|"flow2d_est_induced"|"weight"]``.

B200 mapping per iteration: codd_raft_motion_info (projective transform + log + depth sampling fused) ->
codd_corr_lookup (correlation pyramid never materialised) -> update block convolutions (the four head
stems run as one wide convolution, z|r gates as one) -> codd_se3_gn_step.  The all-pairs volume of the
reference (299 MB / sample at 576x960, blocks/corr.py:28-46) does not exist here.
"""
import torch
import torch.nn as nn

from .. import ops
from ..registry import MODELS, build_backbone
from ._net import NetWeights, conv, conv_cat
from .extractor import BasicEncoder


class GradientClip(nn.Module):
    """Identity in the forward pass (raft3d.py:22-40)."""

    def forward(self, x):
        return x


class ConvGRU(nn.Module):
    """blocks/gru.py:10-35."""

    def __init__(self, hidden_dim=128, input_dim=192 + 128, dilation=4):
        super().__init__()
        self.hidden_dim = hidden_dim
        for g in "zrq":
            setattr(self, f"conv{g}1", nn.Conv2d(hidden_dim, hidden_dim, 3, padding=1))
            setattr(self, f"conv{g}2", nn.Conv2d(hidden_dim, hidden_dim, 3, dilation=dilation, padding=dilation))

    def run(self, pw, h, inp_sum):
        """inp_sum [N,384,h,w] = sum of the inputs' (z|r|q) thirds (gru.py:23-28)."""
        hd = self.hidden_dim
        n, _, hh, ww = h.shape
        # z | r: two 128->256 convolutions (dilation 1, then dilation 4 accumulating) + sigmoid
        zr = conv_cat(pw, [self.convz1, self.convr1], h, ops.ACT_NONE, residual=inp_sum[:, :2 * hd])
        zr = conv_cat(pw, [self.convz2, self.convr2], h, ops.ACT_SIGMOID, residual=zr)
        rh = ops.eltwise(ops.EW_MUL, zr[:, hd:], h)
        q = conv(pw, self.convq1, rh, ops.ACT_NONE, residual=inp_sum[:, 2 * hd:])
        q = conv(pw, self.convq2, rh, ops.ACT_TANH, residual=q)
        return ops.eltwise(ops.EW_GRU, zr[:, :hd], h, q)


class BasicUpdateBlock(nn.Module):
    """raft3d.py:43-106."""

    def __init__(self, hidden_dim=128, input_dim=128):
        super().__init__()
        self.gru = ConvGRU(hidden_dim)
        self.corr_enc = nn.Sequential(nn.Conv2d(196, 256, 3, padding=1), nn.ReLU(inplace=True),
                                      nn.Conv2d(256, 256, 3, padding=1), nn.ReLU(inplace=True),
                                      nn.Conv2d(256, 3 * 128, 1, padding=0))
        self.flow_enc = nn.Sequential(nn.Conv2d(9, 128, 7, padding=3), nn.ReLU(inplace=True),
                                      nn.Conv2d(128, 3 * 128, 1, padding=0))
        self.ae = nn.Sequential(nn.Conv2d(128, 256, 3, padding=1), nn.ReLU(inplace=True),
                                nn.Conv2d(256, 32, 1, padding=0), GradientClip())
        self.delta = nn.Sequential(nn.Conv2d(128, 256, 3, padding=1), nn.ReLU(inplace=True),
                                   nn.Conv2d(256, 3, 1, padding=0), GradientClip())
        self.weight = nn.Sequential(nn.Conv2d(128, 256, 3, padding=1), nn.ReLU(inplace=True),
                                    nn.Conv2d(256, 3, 1, padding=0), nn.Sigmoid(), GradientClip())
        self.mask = nn.Sequential(nn.Conv2d(128, 256, 3, padding=1), nn.ReLU(inplace=True),
                                  nn.Conv2d(256, 64 * 9, 1, padding=0), GradientClip())

    def run(self, pw, net, inp, corr, motion_info, want_mask=True):
        """motion_info is the clamped [flow, 10*twist, 10*dz] tensor (the reference's call passes dz / twist
        swapped relative to the signature, raft3d.py:92-94 vs :238-240; the effective order is kept)."""
        # i_all = inp + corr_enc(corr) + flow_enc(motion_info): the three GRU inputs summed in the epilogues
        c = conv(pw, self.corr_enc[0], corr, ops.ACT_RELU)
        c = conv(pw, self.corr_enc[2], c, ops.ACT_RELU)
        s = conv(pw, self.corr_enc[4], c, ops.ACT_NONE, residual=inp)
        m = conv(pw, self.flow_enc[0], motion_info, ops.ACT_RELU)
        s = conv(pw, self.flow_enc[2], m, ops.ACT_NONE, residual=s)
        net = self.gru.run(pw, net, s)
        heads = [self.ae, self.delta, self.weight] + ([self.mask] if want_mask else [])
        stem = conv_cat(pw, [hd[0] for hd in heads], net, ops.ACT_RELU)
        ae = conv(pw, self.ae[2], stem[:, 0:256])
        delta = conv(pw, self.delta[2], stem[:, 256:512])
        weight = conv(pw, self.weight[2], stem[:, 512:768], ops.ACT_SIGMOID)
        mask = conv(pw, self.mask[2], stem[:, 768:1024]) if want_mask else None
        return net, mask, ae, delta, weight

    def forward(self, net, inp, corr, flow, twist, dz, upsample=True):
        """Reference signature (raft3d.py:92): tensors in, NCHW semantics."""
        info = torch.cat([flow, 10 * dz, 10 * twist], dim=-1).clamp(-50.0, 50.0).permute(0, 3, 1, 2)
        pw = self.__dict__.setdefault("_pw", NetWeights())
        return self.run(pw, ops.to_nhwc(net), ops.to_nhwc(inp), ops.to_nhwc(corr), ops.to_nhwc(info.contiguous()))


class ResizeConcatConv(nn.Module):
    """raft3d.py:109-137: bilinear (align_corners=True) resize of every map to inputs[1]'s size, concat, 1x1 conv
    (no bias) + ReLU.  The resized maps are written straight into the concat buffer."""

    def __init__(self, in_channels, out_channels=32):
        super().__init__()
        assert isinstance(in_channels, (list, tuple))
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.convs = nn.Sequential(nn.Conv2d(sum(in_channels), out_channels, kernel_size=1, padding=0, stride=1, bias=False),
                                   nn.ReLU(inplace=True))

    def run(self, pw, inputs):
        assert len(inputs) == len(self.in_channels)
        n, _, h, w = inputs[1].shape
        ctot = sum(self.in_channels)
        ld = (ctot + 3) // 4 * 4
        cat = ops.empty_nhwc(n, ctot, h, w, inputs[1].device, ld=ld)
        c0 = 0
        for x, c in zip(inputs, self.in_channels):
            ops.resize_bilinear(x, (h, w), True, out=cat[:, c0:c0 + c])
            c0 += c
        return conv(pw, self.convs[0], cat, ops.ACT_RELU)

    def forward(self, inputs):
        pw = self.__dict__.setdefault("_pw", NetWeights())
        return self.run(pw, [ops.to_nhwc(x) for x in inputs])


@MODELS.register_module(force=True)
class RAFT3D(nn.Module):
    def __init__(self, cnet_cfg=None):
        super().__init__()
        self.hidden_dim = hdim = 128
        self.context_dim = 128
        self.corr_levels = 4
        self.corr_radius = 3
        self.fnet = BasicEncoder(output_dim=128, norm_fn="instance")
        if cnet_cfg is None:
            raise NotImplementedError("RAFT3D needs cnet_cfg (the reference's FPN fallback is undefined, raft3d.py:153)")
        self.cnet = nn.Sequential(build_backbone(cnet_cfg),
                                  ResizeConcatConv(cnet_cfg["extra"]["stage4"]["num_channels"], 128 * 4))
        self.update_block = BasicUpdateBlock(hidden_dim=hdim)
        self._pw = NetWeights()

    def context(self, image):
        return self.cnet[1].run(self._pw, self.cnet[0](image))

    def forward(self, image_curr, depth_prev, depth_curr, intrinsics, state, outputs, iters=12, train_mode=False):
        if "memory" not in state:
            state["raft_feat"] = self.fnet(image_curr)
            state["raft_netinp"] = self.context(image_curr)
            return
        pw = self._pw
        fmap_prev, net_inp = ops.to_nhwc(state["raft_feat"]), ops.to_nhwc(state["raft_netinp"])
        n, _, H, W = image_curr.shape
        fmap_curr = self.fnet(image_curr)
        pyramid = ops.corr_pyramid(fmap_curr, self.corr_levels)
        net = ops.eltwise(ops.EW_ACT, net_inp[:, :128], act=ops.ACT_TANH)
        inp = ops.eltwise(ops.EW_ACT, net_inp[:, 128:], act=ops.ACT_RELU)

        intr = intrinsics.to(device=image_curr.device, dtype=torch.float32).contiguous()
        intr8 = (intr / 8.0).contiguous()
        depth_prev = depth_prev.contiguous()
        depth1_r8 = ops.subsample(depth_prev.unsqueeze(-1), 3, 8).squeeze(-1)          # depth_prev[:, 3::8, 3::8]
        depth2inv_r8 = ops.subsample(depth_curr.contiguous().unsqueeze(-1), 3, 8, recip=True).squeeze(-1)
        h8, w8 = H // 8, W // 8
        Ts = torch.zeros((n, h8, w8, 7), device=image_curr.device)      # SE3.Identity: t = 0, q = (0,0,0,1)
        Ts[..., 6] = 1.0

        mask = weight = None
        for it in range(iters):
            xyz, info = ops.raft_motion_info(Ts, depth1_r8, depth2inv_r8, intr8)
            corr = ops.corr_lookup(fmap_prev, pyramid, xyz, self.corr_radius)
            # the convex-upsampling mask is consumed after the last iteration only (the reference's eval-time
            # train_mode branch, raft3d.py:249-265, fills outputs nothing reads: SURVEY.md appendix D.1)
            net, mask, ae, delta, weight = self.update_block.run(pw, net, inp, corr, info, want_mask=it == iters - 1)
            target = ops.eltwise(ops.EW_ADD_ACT, xyz.permute(0, 3, 1, 2), delta)
            Ts = ops.se3_gn_step(Ts, ae, target, weight, depth1_r8, intr8)

        Ts_up, flow = ops.se3_upsample_flow(Ts, mask, depth_prev, intr)
        outputs["Ts"] = Ts_up
        outputs["flow2d_est_induced"] = flow
        outputs["weight"] = ops.cvx_upsample(weight.permute(0, 2, 3, 1), mask).permute(0, 3, 1, 2)
        state["raft_feat"] = fmap_curr
        state["raft_netinp"] = self.context(image_curr)
