"""ConsistentOnlineDynamicDepth — drop-in for model/codd.py:21-126, 269-398 (inference path).

Keeps the reference's construction (``stereo`` / ``motion`` / ``fusion`` config dicts built
through the registry), call signature ``model(return_loss=False, rescale=True, evaluate=False,
img=[T[B,MF,3,H,W]], img_metas=[[dict]], r_img=[T[B,MF,3,H,W]])`` and result
``[T[B,MF,img_h,img_w]]``.  Training (``return_loss=True``), metric evaluation
(``evaluate=True``) and result dumping are host-side bookkeeping outside the accelerated path
 and raise ``NotImplementedError``.
"""
from collections import OrderedDict

import torch
import torch.nn as nn

from .registry import ESTIMATORS, MODELS


@ESTIMATORS.register_module(force=True)
class ConsistentOnlineDynamicDepth(nn.Module):
    def __init__(self, stereo=None, motion=None, fusion=None, train_cfg=None, test_cfg=None, init_cfg=None,
                 **kwargs):
        super().__init__()
        self.fp16_enabled = False
        self.train_cfg = train_cfg
        self.test_cfg = test_cfg
        self.init_cfg = init_cfg
        self.build_model(stereo, motion, fusion)

    def build_model(self, stereo, motion, fusion):
        assert stereo is not None
        self.stereo = MODELS.build(stereo)
        self.motion = MODELS.build(motion) if motion is not None else None
        self.fusion = MODELS.build(fusion) if fusion is not None else None

    def _frozen(self, key):
        return (self.train_cfg is not None) and bool(self.train_cfg.get(key, False))

    def freeze_fusion(self):
        return self._frozen("freeze_fusion")

    def freeze_motion(self):
        return self._frozen("freeze_motion")

    def freeze_stereo(self):
        return self._frozen("freeze_stereo")

    def consistent_online_depth_estimation(self, left_img, right_img, img_metas, state):
        """stereo -> motion -> fusion query -> fusion update for one frame (codd.py:80-126)."""
        with torch.no_grad():
            outputs = self.stereo.stereo_matching(left_img, right_img, img_metas, state)
            if self.motion is not None:
                # reference quirk: `not a & b` parses as `not (a & b)` => train_mode is True in eval
                self.motion(state, outputs, img_metas=img_metas,
                            train_mode=not (self.freeze_motion() & self.training))
            if self.fusion is not None:
                self.fusion.memory_query(outputs, state, img_metas=img_metas)
                self.fusion.memory_update(outputs, state, img_metas=img_metas)
        return outputs

    def forward(self, img, img_metas, return_loss=True, **kwargs):
        if return_loss:
            raise NotImplementedError("codd_b200 is a forward/inference-only build (training: out of scope)")
        return self.forward_test(img, img_metas, **kwargs)

    def forward_test(self, img, img_metas, r_img=None, **kwargs):
        for var, name in [(img, "img"), (img_metas, "img_metas")]:
            if not isinstance(var, list):
                raise TypeError(f"{name} must be a list, but got {type(var)}")
        img = img[0]
        r_img = r_img[0] if r_img is not None else r_img
        with torch.no_grad():
            pred = self.inference(img, r_img, img_metas[0], **kwargs)
        return [pred]

    def inference(self, img, r_img, img_meta, reciprocal=False, evaluate=True, **kwargs):
        if evaluate:
            raise NotImplementedError("metric evaluation is host-side bookkeeping outside this build; "
                                      "call with evaluate=False and score the returned disparities")
        self.reset_inference_state()
        l_img_list = torch.unbind(img, dim=1)
        r_img_list = torch.unbind(r_img, dim=1)
        img_h, img_w = img_meta[0]["img_shape"][:2]
        outputs = []
        for l_img, r_img_t in zip(l_img_list, r_img_list):
            output = self.consistent_online_depth_estimation(l_img, r_img_t, img_meta, self.inference_state)
            pred_disp = output["pred_disp"]
            if reciprocal:
                pred_disp = img_meta[0]["calib"] / pred_disp
            self.inference_state["pred_disp"].append(pred_disp)
            outputs.append(pred_disp[:, :, :img_h, :img_w])
        outputs = torch.cat(outputs, dim=1)
        assert len(outputs.shape) == 4, "Output shape is wrong"
        return outputs

    def reset_inference_state(self):
        self.inference_state = OrderedDict(pred_disp=[])

    def train(self, mode=True):
        """Reference quirk kept: ``train()`` / ``eval()`` return None (codd.py:601-630)."""
        self.training = mode
        for m in self.children():
            m.train(mode)
