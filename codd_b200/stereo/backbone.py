"""HITUNet — drop-in for model/stereo/hitnet/backbone.py:42-88 (registry name ``HITUNet``).

Same parameter tree (conv1, down1..4, up4..1, merge4..1).  ``forward`` returns the reference's
5-level pyramid ``[1/16 x32, 1/8 x24, 1/4 x24, 1/2 x16, 1/1 x16]`` as logical-NCHW tensors in
NHWC memory.  ``forward_pair`` pushes the left and right images through one batched pass (the
weights are shared, so the reference's two ``extract_feat`` calls become one set of launches).
"""
import torch
import torch.nn as nn

from .. import ops
from ..lib import ACT_LEAKY
from ..registry import BACKBONES
from ._params import PackedWeights, run_conv, run_conv_pair


def _conv_down(inp, oup):
    return nn.Sequential(
        nn.Conv2d(inp, oup, 4, stride=2, padding=1), nn.LeakyReLU(0.2, inplace=True),
        nn.Conv2d(oup, oup, 3, stride=1, padding=1), nn.LeakyReLU(0.2, inplace=True))


def _conv_up(inp, oup):
    return nn.Sequential(nn.ConvTranspose2d(inp, oup, 2, stride=2, padding=0), nn.LeakyReLU(0.2, inplace=True))


def _conv_merge(inp, oup):
    return nn.Sequential(
        nn.Conv2d(inp, oup, 1, stride=1, padding=0), nn.LeakyReLU(0.2, inplace=True),
        nn.Conv2d(oup, oup, 3, stride=1, padding=1), nn.LeakyReLU(0.2, inplace=True),
        nn.Conv2d(oup, oup, 3, stride=1, padding=1), nn.LeakyReLU(0.2, inplace=True))


@BACKBONES.register_module(force=True)
class HITUNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv2d(3, 16, 3, stride=1, padding=1), nn.LeakyReLU(0.2, inplace=True))
        self.down1 = _conv_down(16, 16)
        self.down2 = _conv_down(16, 24)
        self.down3 = _conv_down(24, 24)
        self.down4 = nn.Sequential(
            _conv_down(24, 32),
            nn.Conv2d(32, 32, 3, stride=1, padding=1), nn.LeakyReLU(0.2, inplace=True),
            nn.Conv2d(32, 32, 3, stride=1, padding=1), nn.LeakyReLU(0.2, inplace=True))
        self.up4 = _conv_up(32, 24)
        self.up3 = _conv_up(24, 24)
        self.up2 = _conv_up(24, 16)
        self.up1 = _conv_up(16, 16)
        self.merge4 = _conv_merge(24 + 24, 24)
        self.merge3 = _conv_merge(24 + 24, 24)
        self.merge2 = _conv_merge(16 + 16, 16)
        self.merge1 = _conv_merge(16 + 16, 16)
        self._pw = PackedWeights()

    # -- building blocks -------------------------------------------------------------------
    def _c(self, conv, x, x2=None):
        return run_conv(self._pw, conv, x, ACT_LEAKY, x2=x2)

    def _down(self, seq, x):
        return self._c(seq[2], self._c(seq[0], x))

    def _up(self, seq, x):
        wp, b = self._pw.deconv(seq[0])
        return ops.deconv2x2(x, wp, b, seq[0].out_channels, ACT_LEAKY)

    def _merge(self, seq, skip, up):
        return run_conv_pair(self._pw, seq[2], ACT_LEAKY, seq[4], ACT_LEAKY, self._c(seq[0], skip, up))

    def _up_merge(self, up_seq, merge_seq, skip, coarse):
        """conv_up followed by conv_merge (backbone.py:17-32,75-88); the deconvolution and the 1x1 that consumes it run as
        one kernel that never writes the up-sampled tensor."""
        cu, co = up_seq[0].out_channels, merge_seq[0].out_channels
        if ops.upmerge_eligible(coarse, skip, cu, co):
            wu, bu = self._pw.deconv(up_seq[0])
            wm, bm = self._pw.conv(merge_seq[0])
            x = ops.upmerge(coarse, skip, wu, bu, cu, wm, bm, co)
            return run_conv_pair(self._pw, merge_seq[2], ACT_LEAKY, merge_seq[4], ACT_LEAKY, x)
        return self._merge(merge_seq, skip, self._up(up_seq, coarse))

    def _features(self, left, right):
        wp, b = self._pw.conv(self.conv1[0])
        x0 = ops.conv3x3_image(left, right, wp, b, 16)
        x1 = self._down(self.down1, x0)
        x2 = self._down(self.down2, x1)
        x3 = self._down(self.down3, x2)
        x4 = self._down(self.down4[0], x3)
        x4 = self._c(self.down4[3], self._c(self.down4[1], x4))
        u4 = self._up_merge(self.up4, self.merge4, x3, x4)
        u3 = self._up_merge(self.up3, self.merge3, x2, u4)
        u2 = self._up_merge(self.up2, self.merge2, x1, u3)
        u1 = self._up_merge(self.up1, self.merge1, x0, u2)
        return [x4, u4, u3, u2, u1]

    # -- public ------------------------------------------------------------------------------
    def forward(self, x):
        return self._features(x, None)

    def forward_pair(self, left, right):
        n = left.shape[0]
        feats = self._features(left, right)
        return [f[:n] for f in feats], [f[n:] for f in feats]
