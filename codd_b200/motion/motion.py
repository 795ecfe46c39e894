"""Motion — drop-in for model/motion/motion.py:48-209 (registry name ``Motion``), inference path.

Same constructor (``raft3d, ds_scale, iters, loss``) and ``forward(state, outputs, img_metas, train_mode)``
contract: the first frame only primes ``state["raft_feat"/"raft_netinp"]``; later frames run RAFT3D and replace
``state["memory"]`` (the 3-tuple written by ``Fusion.memory_update``) with the 5-tuple
``[img_warp, feat_warp, confidence_warp, disp_warp, flow_warp]`` aligned to the current frame.

The pytorch3d point renderer of the reference (transform_and_project, motion.py:82-130) is replaced by
codd_splat_warp (z-sorted top-8 alpha compositing); disparity <-> depth conversion and the 1/4 sub-sampling are
codd_disp_to_depth / codd_subsample_nhwc.
"""
import torch
import torch.nn as nn

from .. import ops
from ..registry import MODELS, build_loss

BF_DEFAULT = 1050 * 0.2  # baseline * focal length (motion.py:45)


@MODELS.register_module(force=True)
class Motion(nn.Module):
    def __init__(self, raft3d=None, ds_scale=4, iters=16, loss=None):
        super().__init__()
        self.ds_scale = ds_scale
        self.iters = iters
        self.raft3d = MODELS.build(raft3d)
        self.loss = build_loss(loss) if loss is not None else None
        n_parameters = sum(p.numel() for n, p in self.named_parameters())
        print("PARAM STATUS: total number of parameters %.3fM in motion network" % (n_parameters / 1000 ** 2))

    def transform_and_project(self, Ts, depth, feat, intrinsics, radius, bf=0.0, want_disp=False):
        """Ts [N,H,W,7], depth [N,H,W], feat [N,C,H,W] -> (aligned feature, z-buffer[, disparity])."""
        out, zbuf, disp = ops.splat_warp(Ts, depth, intrinsics, ops.to_nhwc(feat), radius, bf=bf, want_disp=want_disp)
        return (out, zbuf, disp) if want_disp else (out, zbuf)

    def forward(self, state, outputs, img_metas, train_mode=False, **kwargs):
        img_curr = outputs["left_img"]
        if "memory" not in state:
            self.raft3d(img_curr, None, None, None, state, outputs, train_mode=train_mode)
            return
        dev = outputs["pred_disp"].device
        B = outputs["pred_disp"].shape[0]
        intrinsics = torch.tensor(img_metas[0]["intrinsics"])
        fx = intrinsics[0]
        depth_scale = BF_DEFAULT / fx                       # same rounding sequence as motion.py:153-157
        bf = float((depth_scale * fx).float())
        intrinsics = intrinsics.float().to(dev).unsqueeze(0).expand(B, -1).contiguous()

        img_prev, feat_prev, disp_prev = state["memory"]
        disp_curr = outputs["pred_disp"]
        depth_prev = ops.disp_to_depth(disp_prev.reshape(B, *disp_prev.shape[-2:]), bf)
        depth_curr = ops.disp_to_depth(disp_curr.reshape(B, *disp_curr.shape[-2:]), bf)

        self.raft3d(img_curr, depth_prev, depth_curr, intrinsics, state, outputs, iters=self.iters,
                    train_mode=train_mode)
        Ts = outputs["Ts"]

        # full-resolution warp of [img_prev | induced flow | confidence] (motion.py:181-193)
        n, _, H, W = img_curr.shape
        to_proj = ops.empty_nhwc(n, 9, H, W, dev, ld=12)
        ops.copy_to_nhwc(img_prev, to_proj[:, 0:3])
        ops.copy_to_nhwc(outputs["flow2d_est_induced"].permute(0, 3, 1, 2), to_proj[:, 3:6])
        ops.copy_to_nhwc(outputs["weight"], to_proj[:, 6:9])
        warped, _, disp_warp = self.transform_and_project(Ts, depth_prev, to_proj, intrinsics, 2.0, bf=bf, want_disp=True)
        img_warp, flow_warp, confidence_warp = warped[:, :3], warped[:, 3:6], warped[:, 6:]

        # low-resolution feature warp (motion.py:195-203)
        ds = self.ds_scale
        Ts_lr = ops.subsample(Ts, ds // 2 - 1, ds)
        depth_lr = ops.subsample(depth_prev.unsqueeze(-1), ds // 2 - 1, ds).squeeze(-1)
        feat_warp, _ = self.transform_and_project(Ts_lr, depth_lr, feat_prev, (intrinsics / ds).contiguous(), 4.0)

        state["memory"] = [img_warp, feat_warp, confidence_warp, disp_warp, flow_warp]

    def losses(self, *args, **kwargs):
        raise NotImplementedError("codd_b200 is a forward-only build (training losses: out of scope, DESIGN.md)")

    def freeze(self):
        self.eval()
        if self.loss is not None:
            self.loss.eval()
        for param in self.parameters():
            param.requires_grad = False
