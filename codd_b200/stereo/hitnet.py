"""HITNetMF — drop-in for model/stereo/hitnet/hitnet.py:13-122 (registry name ``HITNetMF``).

Same constructor (backbone / initialization / propagation / loss config dicts built through
the registry), same ``stereo_matching`` output dict.  ``pred_disp`` is a plain contiguous
[N,1,H,W] tensor; ``left_feat`` / ``right_feat`` are logical [N,24,H/4,W/4] tensors in NHWC
(channels_last) memory.
"""
import torch
import torch.nn as nn

from .. import ops
from ..registry import ESTIMATORS, MODELS, build_backbone, build_loss


@ESTIMATORS.register_module(force=True)
class HITNetMF(nn.Module):
    def __init__(self, backbone, initialization, propagation, loss=None):
        super().__init__()
        self.backbone = build_backbone(backbone)
        self.tile_init = MODELS.build(initialization)
        self.tile_update = MODELS.build(propagation)
        self.freezed = False
        self.loss = build_loss(loss) if loss is not None else None
        n_parameters = sum(p.numel() for n, p in self.named_parameters())
        print("PARAM STATUS: total number of parameters %.3fM in stereo network" % (n_parameters / 1000 ** 2))

    def extract_feat(self, img):
        return self.backbone(img)

    def losses(self, *args, **kwargs):
        raise NotImplementedError("codd_b200 is a forward-only build (training losses: out of scope, DESIGN.md)")

    def stereo_matching(self, left_img, right_img, img_metas=None, state=None):
        if self.training and not self.freezed:
            raise NotImplementedError("codd_b200.HITNetMF: training forward is out of scope (DESIGN.md)")
        if hasattr(self.backbone, "forward_pair"):
            left_fea_pyramid, right_fea_pyramid = self.backbone.forward_pair(left_img, right_img)
        else:
            left_fea_pyramid = self.extract_feat(left_img)
            right_fea_pyramid = self.extract_feat(right_img)
        init_cv_pyramid, init_tile_pyramid = self.tile_init(left_fea_pyramid, right_fea_pyramid)
        pred = self.tile_update(left_fea_pyramid, right_fea_pyramid, init_tile_pyramid)
        outputs = dict(pred_disp=pred, left_feat=left_fea_pyramid[2], right_feat=right_fea_pyramid[2])
        outputs["left_img"] = left_img
        if len(outputs["pred_disp"].shape) == 3:
            outputs["pred_disp"] = outputs["pred_disp"].unsqueeze(1)
        return outputs

    def freeze(self):
        for m in (self.tile_update, self.tile_init, self.backbone):
            m.eval()
            for param in m.parameters():
                param.requires_grad = False
        if self.loss is not None:
            self.loss.eval()
            for param in self.loss.parameters():
                param.requires_grad = False
        self.freezed = True
