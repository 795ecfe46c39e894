"""TileInitialization — drop-in for model/stereo/hitnet/initialization.py:48-230.

Same constructor (``max_disp``, ``fea_c``) and parameter tree (tile_conv{1x..16x},
tile_fea_dscrpt{16x..1x}).  ``forward(fea_l_pyramid, fea_r_pyramid)`` returns
``[init_cv_pyramid, init_hypo_pyramid]`` like the reference; the cost volumes are only
materialised when asked for (``materialize_cv`` — the reference's eval path computes and
drops them, hitnet.py:89-94), otherwise the fused arg-min kernel runs and the first list
holds ``None`` entries.
"""
import torch
import torch.nn as nn

from .. import ops
from ..lib import ACT_LEAKY
from ..registry import MODELS
from ._params import PackedWeights


def _tile_conv(cin):
    return nn.Sequential(
        nn.Conv2d(cin, 16, 4, 4, 0), nn.LeakyReLU(0.2, inplace=True),
        nn.Conv2d(16, 16, 1, 1, 0), nn.LeakyReLU(0.2, inplace=True))


def _dscrpt(cin):
    return nn.Sequential(nn.Conv2d(cin, 13, 1), nn.LeakyReLU(0.2, inplace=True))


@MODELS.register_module(force=True)
class TileInitialization(nn.Module):
    def __init__(self, max_disp, fea_c=[16, 16, 24, 24, 32]):
        super().__init__()
        self.maxdisp = max_disp
        fea_c1x, fea_c2x, fea_c4x, fea_c8x, fea_c16x = fea_c
        self.pad = nn.ZeroPad2d((0, 3, 0, 0))
        self.tile_conv1x = _tile_conv(fea_c1x)
        self.tile_conv2x = _tile_conv(fea_c2x)
        self.tile_conv4x = _tile_conv(fea_c4x)
        self.tile_conv8x = _tile_conv(fea_c8x)
        self.tile_conv16x = _tile_conv(fea_c16x)
        self.tile_fea_dscrpt16x = _dscrpt(17)
        self.tile_fea_dscrpt8x = _dscrpt(17)
        self.tile_fea_dscrpt4x = _dscrpt(33)
        self.tile_fea_dscrpt2x = _dscrpt(25)
        self.tile_fea_dscrpt1x = _dscrpt(25)
        self.materialize_cv = None  # None: follow self.training
        self._pw = PackedWeights()

    def _levels(self):
        # coarse -> fine, matching the pyramid order the backbone emits
        return [(self.tile_conv16x, self.tile_fea_dscrpt16x, 16), (self.tile_conv8x, self.tile_fea_dscrpt8x, 8),
                (self.tile_conv4x, self.tile_fea_dscrpt4x, 4), (self.tile_conv2x, self.tile_fea_dscrpt2x, 2),
                (self.tile_conv1x, self.tile_fea_dscrpt1x, 1)]

    def _tile_pair(self, seq, fl, fr):
        """initialization.py:119-124: left 4x4/s4; right the same weights at stride (4,1) over the
        input zero-padded by 3 columns on the right.  One fused kernel per side (K2), planar out."""
        w1, b1 = self._pw.raw(seq[2])
        if ops.tile_features_tc_eligible(fl):
            # the 4x4 conv as a tcgen05 implicit GEMM (csrc/conv_tc_s2.cu)
            ws, b0 = self._pw.conv_tc4(seq[0])
            return (ops.tile_features_tc(fl, ws, b0, w1, b1, right=False),
                    ops.tile_features_tc(fr, ws, b0, w1, b1, right=True))
        w0, b0 = self._pw.conv(seq[0])
        return ops.tile_features(fl, w0, b0, w1, b1, right=False), ops.tile_features(fr, w0, b0, w1, b1, right=True)

    def tile_features(self, fea_l, fea_r):
        fea_l = [ops.to_nhwc(f) for f in fea_l]
        fea_r = [ops.to_nhwc(f) for f in fea_r]
        return [list(self._tile_pair(seq, fea_l[k], fea_r[k])) for k, (seq, _, _) in enumerate(self._levels())]

    def tile_hypothesis_pyramid(self, tile_feature_pyramid, fea_l_pyramid):
        want_cv = self.training if self.materialize_cv is None else self.materialize_cv
        cvs, hyps = [], []
        levels = self._levels()
        # K1 for the five levels in one launch (the coarse levels fill the tail of the finest one)
        k1 = ops.cost_volume_pyramid(tile_feature_pyramid, [self.maxdisp // div for _, _, div in levels], want_cv=want_cv)
        for k, (_, dsc, div) in enumerate(levels):
            tl, tr = tile_feature_pyramid[k]
            cv, cost, disp = k1[k]
            # descriptor input: tile features at 16x / 8x, backbone pyramid [0..2] below (Eq. 4)
            feat = tl if k < 2 else ops.to_nhwc(fea_l_pyramid[k - 2])
            w, b = self._pw.raw(dsc[0])
            hyps.append(ops.tile_hyp_init(cost, disp, feat, w, b))
            cvs.append(cv)
        return [cvs, hyps]

    def forward(self, fea_l_pyramid, fea_r_pyramid):
        tiles = self.tile_features(fea_l_pyramid, fea_r_pyramid)
        return self.tile_hypothesis_pyramid(tiles, fea_l_pyramid)
