"""ConsistentOnlineDynamicDepth — drop-in for model/codd.py:21-126, 269-398 (inference path).

Keeps the reference's construction (``stereo`` / ``motion`` / ``fusion`` config dicts built
through the registry), call signature ``model(return_loss=False, rescale=True, evaluate=False,
img=[T[B,MF,3,H,W]], img_metas=[[dict]], r_img=[T[B,MF,3,H,W]])`` and result
``[T[B,MF,img_h,img_w]]``; ``evaluate=True`` returns the reference's meters, accumulated on the GPU
(SURVEY.md §8f N3).  Training (``return_loss=True``) and result dumping are outside the accelerated
path: training raises ``NotImplementedError``.
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
        """codd.py:290-398.  evaluate=False: the disparity maps [B,MF,img_h,img_w].  evaluate=True: the meters of
        calc_metric / collect_metric (codd.py:435-575, utils/misc.py:62-77) as {name: tensor([value])}, accumulated on
        the GPU by `codd_b200.metrics.SequenceMetrics` (no per-frame host synchronisation) from the ground truth passed
        as keyword lists (gt_disp, gt_flow, gt_disp_change, gt_flow_occ, gt_disp2, gt_disp_occ: utils/misc.py:93-134)."""
        self.reset_inference_state()
        l_img_list = torch.unbind(img, dim=1)
        r_img_list = torch.unbind(r_img, dim=1)
        img_h, img_w = img_meta[0]["img_shape"][:2]
        gts = metrics = None
        if evaluate:
            gts = {k: (None if kwargs.get(k) is None else torch.unbind(kwargs[k][0], dim=1))
                   for k in ("gt_disp", "gt_flow", "gt_disp_change", "gt_flow_occ", "gt_disp2", "gt_disp_occ")}
            assert gts["gt_disp"] is not None, "No ground truth provided"
            from .metrics import SequenceMetrics
            metrics = SequenceMetrics(img_meta[0]["disp_range"], max_frames=len(l_img_list), device=img.device)
            self.inference_state.update(gt_disp=[], gt_flow_occ=[], gt_disp_change=[])
        outputs = []
        for idx, (l_img, r_img_t) in enumerate(zip(l_img_list, r_img_list)):
            output = self.consistent_online_depth_estimation(l_img, r_img_t, img_meta, self.inference_state)
            pred_disp = output["pred_disp"]
            if reciprocal:
                pred_disp = img_meta[0]["calib"] / pred_disp
            self.inference_state["pred_disp"].append(pred_disp)
            outputs.append(pred_disp[:, :, :img_h, :img_w])
            if evaluate:
                self._evaluate_frame(metrics, idx, gts, pred_disp, img_h, img_w, output.get("Ts", None), img_meta[0])
        if evaluate:
            return {k: torch.tensor([v]) for k, v in metrics.collect().items()}
        outputs = torch.cat(outputs, dim=1)
        assert len(outputs.shape) == 4, "Output shape is wrong"
        return outputs

    def _evaluate_frame(self, metrics, idx, gts, pred_disp, img_h, img_w, Ts, meta):
        """The ground-truth bookkeeping of codd.py:313-355 for frame `idx`, then one SequenceMetrics.update (which keeps
        the previous frame's tensors itself: inference_state[...][-2] in the reference)."""
        st = self.inference_state

        def crop(name):
            return None if gts[name] is None else gts[name][idx][:, :, :img_h, :img_w]

        gt_disp, gt_flow, gt_disp2 = crop("gt_disp"), crop("gt_flow"), crop("gt_disp2")
        st["gt_disp"].append(gt_disp)
        if gts["gt_disp_change"] is not None:
            st["gt_disp_change"].append(crop("gt_disp_change"))
        if gts["gt_flow_occ"] is not None:
            st["gt_flow_occ"].append((gts["gt_flow_occ"][idx] > 0)[:, :, :img_h, :img_w])     # True = occluded
            if gts["gt_disp_change"] is None and idx > 0:                                     # codd.py:331-340
                from . import ops
                change, _ = ops.gt_disp_change(gts["gt_flow"][idx - 1][:, :, :img_h, :img_w], gt_disp,
                                               st["gt_disp"][idx - 1], st["gt_flow_occ"][idx - 1])
                st["gt_disp_change"].append(change)
        if gt_disp2 is not None and gts["gt_disp_change"] is None:                           # codd.py:343-349
            change = gt_disp2 - gt_disp
            change[gt_disp2 <= 0.0] = 1050 * 0.2
            change[gt_disp <= 0.0] = 1050 * 0.2
            st["gt_disp_change"].append(change)
        seg = None if gts["gt_disp_occ"] is None else (gts["gt_disp_occ"][idx] <= 0)[:, :, :img_h, :img_w].float()
        motion = {}
        if idx > 0 and Ts is not None and len(st["gt_disp_change"]) > 0:                      # codd.py:519-540
            if len(st["gt_flow_occ"]) > 0:
                motion = dict(gt_disp_change=st["gt_disp_change"][-1], gt_flow_occ_prev=st["gt_flow_occ"][-2])
            elif len(st["gt_disp_change"]) > 1:
                motion = dict(gt_disp_change=st["gt_disp_change"][-2])
            if motion:
                intr = torch.tensor(meta["intrinsics"], dtype=torch.float32, device=pred_disp.device)
                motion.update(Ts=Ts[:, :img_h, :img_w], intrinsics=intr.unsqueeze(0).expand(pred_disp.shape[0], -1))
        metrics.update(pred_disp, gt_disp, gt_flow=gt_flow, seg=seg, gt_disp2=gt_disp2, **motion)

    def reset_inference_state(self):
        self.inference_state = OrderedDict(pred_disp=[])

    def train(self, mode=True):
        """Reference quirk kept: ``train()`` / ``eval()`` return None (codd.py:601-630)."""
        self.training = mode
        for m in self.children():
            m.train(mode)
