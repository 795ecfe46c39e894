"""TilePropagation — drop-in for model/stereo/hitnet/propagation.py:336-454.

Same parameter tree (tile_update0..4, tile_update4_1, tile_update5, tile_update6 with their
decrease / conv0 / resblock / lastconv members).  Forward (eval) returns ``final_disp``
[N,1,H,W].  The training-mode output pyramids need autograd through the kernels and are out
of scope of this forward-only build.
"""
import torch
import torch.nn as nn

from .. import ops
from ..lib import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_RELU_CH0
from ..registry import MODELS
from ._params import PackedWeights, run_conv, run_conv_pair


def _convbn(cin, cout, k, stride, pad, dilation):
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=k, stride=stride,
                                   padding=dilation if dilation > 1 else pad, dilation=dilation))


class BasicBlock(nn.Module):
    """Parameter container of the reference ResNet block (propagation.py:103-121)."""
    expansion = 1

    def __init__(self, c1, c2, s, downsample, p, d):
        super().__init__()
        self.conv1 = nn.Sequential(_convbn(c1, c2, 3, s, p, d), nn.LeakyReLU(0.2, inplace=True))
        self.conv2 = _convbn(c2, c2, 3, 1, p, d)
        self.stride = s


def _resblock(c, d=1):
    return nn.Sequential(BasicBlock(c, c, s=1, p=1, downsample=None, d=d), nn.LeakyReLU(0.2, inplace=True))


class _Runner:
    """Kernel-launch helpers shared by the update modules."""

    def _conv(self, conv, x, act, x2=None, residual=None, res_bcast=False, head=None):
        return run_conv(self._pw, conv, x, act, x2=x2, residual=residual, res_bcast=res_bcast, head=head)

    def _res(self, blk, x):
        """Sequential(BasicBlock, LeakyReLU): lrelu(conv2(lrelu(conv1(x))) + x)."""
        bb = blk[0]
        return run_conv_pair(self._pw, bb.conv1[0][0], ACT_LEAKY, bb.conv2[0], ACT_LEAKY, x, residual=x)


class TileUpdate0(nn.Module, _Runner):
    def __init__(self, in_c, out_c, hid_c):
        super().__init__()
        self.decrease = nn.Sequential(nn.Conv2d(64, 16, 1, stride=1, padding=0), nn.LeakyReLU(0.2, inplace=True))
        self.conv0 = nn.Sequential(nn.Conv2d(in_c, hid_c, 1, stride=1, padding=0), nn.LeakyReLU(0.2, inplace=True))
        self.resblock0 = _resblock(32)
        self.resblock1 = _resblock(32)
        self.lastconv = nn.Conv2d(hid_c, out_c, 3, 1, 1)
        self._pw = PackedWeights()

    def forward(self, fea_l, fea_r, current_hypothesis):
        dw, db = self._pw.raw(self.decrease[0])
        cur = ops.to_nhwc(current_hypothesis)
        aug = ops.tile_warp_cost(ops.to_nhwc(fea_l), ops.to_nhwc(fea_r), cur, None, dw, db)
        u = self._conv(self.conv0[0], aug, ACT_LEAKY)
        u = self._res(self.resblock0, u)
        u = self._res(self.resblock1, u)
        # refined = cur + update, ReLU on the disparity channel (propagation.py:170-171)
        return [self._conv(self.lastconv, u, ACT_RELU_CH0, residual=aug[:, :16])]


class TileUpdate(nn.Module, _Runner):
    def __init__(self):
        super().__init__()
        self.decrease = nn.Sequential(nn.Conv2d(64, 16, 1, stride=1, padding=0), nn.LeakyReLU(0.2, inplace=True))
        self.conv0 = nn.Sequential(nn.Conv2d(64, 32, 1, stride=1, padding=0), nn.LeakyReLU(0.2, inplace=True))
        self.resblock0 = _resblock(32)
        self.resblock1 = _resblock(32)
        self.lastconv = nn.Conv2d(32, 34, 3, 1, 1)
        self._pw = PackedWeights()

    def forward(self, fea_l, fea_r, current_hypothesis, prev_hypothesis):
        dw, db = self._pw.raw(self.decrease[0])
        aug = ops.tile_warp_cost(ops.to_nhwc(fea_l), ops.to_nhwc(fea_r), ops.to_nhwc(current_hypothesis),
                                 ops.to_nhwc(prev_hypothesis), dw, db)
        u = self._conv(self.conv0[0], aug, ACT_LEAKY)
        u = self._res(self.resblock0, u)
        u = self._res(self.resblock1, u)
        u = self._conv(self.lastconv, u, ACT_NONE)
        if getattr(self, "keep_aux", False):     # parity tests read the confidences behind the arg-max select
            self.aux = dict(update=u, aug=aug)
        # eval path needs only the selected hypothesis; the two auxiliary tensors of the
        # reference's return list feed the training losses only (propagation.py:241-248)
        return [ops.hyp_select(u, aug)]


class PostTileUpdate(nn.Module, _Runner):
    def __init__(self, in_c, out_c, hid_c, resblk_num):
        super().__init__()
        self.conv1 = nn.Sequential(
            nn.Conv2d(in_c, hid_c, 1, stride=1, padding=0), nn.LeakyReLU(0.2, inplace=True),
            nn.Conv2d(hid_c, hid_c, 3, stride=1, padding=1), nn.LeakyReLU(0.2, inplace=True))
        blocks = nn.ModuleList()
        for i in range(resblk_num):
            blocks.append(_resblock(hid_c, 3 if i == 1 else 1))
        self.resblocks = nn.Sequential(*blocks)
        self.lastconv = nn.Conv2d(hid_c, out_c, kernel_size=3, padding=1)
        self._pw = PackedWeights()

    def forward(self, fea_l, prev_hypothesis):
        prev = ops.to_nhwc(prev_hypothesis)
        x = self._conv(self.conv1[0], ops.to_nhwc(fea_l), ACT_LEAKY, x2=prev)
        x = self._conv(self.conv1[2], x, ACT_LEAKY)
        for blk in self.resblocks:
            x = self._res(blk, x)
        return self._conv(self.lastconv, x, ACT_RELU_CH0, residual=prev)


class FinalTileUpdate(nn.Module, _Runner):
    def __init__(self, in_c, out_c, hid_c, resblk_num):
        super().__init__()
        self.conv1 = nn.Sequential(
            nn.Conv2d(in_c, hid_c, 1, stride=1, padding=0), nn.LeakyReLU(0.2, inplace=True),
            nn.Conv2d(hid_c, hid_c, 3, stride=1, padding=1), nn.LeakyReLU(0.2, inplace=True))
        blocks = nn.ModuleList()
        for _ in range(resblk_num):
            blocks.append(_resblock(hid_c, 1))
        self.resblocks = nn.Sequential(*blocks)
        self.lastconv = nn.Conv2d(hid_c, out_c, kernel_size=3, padding=1)
        self.full_output = False  # True: emit all out_c channels like the reference module
        self._pw = PackedWeights()

    def forward(self, fea_l, prev_hypothesis):
        prev = ops.to_nhwc(prev_hypothesis)
        x = self._conv(self.conv1[0], ops.to_nhwc(fea_l), ACT_LEAKY, x2=prev)
        x = self._conv(self.conv1[2], x, ACT_LEAKY)
        for blk in self.resblocks:
            x = self._res(blk, x)
        # relu(prev[:, 0:1] + update): the single-channel residual broadcasts over the outputs.
        # Inference consumes channel 0 only (propagation.py:372), so only that filter is run.
        return self._conv(self.lastconv, x, ACT_RELU, residual=prev[:, 0:1], res_bcast=True,
                          head=None if self.full_output else 1)


@MODELS.register_module(force=True)
class TilePropagation(nn.Module):
    def __init__(self):
        super().__init__()
        self.tile_update0 = TileUpdate0(32, 16, 32)
        self.tile_update1 = TileUpdate()
        self.tile_update2 = TileUpdate()
        self.tile_update3 = TileUpdate()
        self.tile_update4 = TileUpdate()
        self.tile_update4_1 = PostTileUpdate(40, 16, 32, 4)
        self.tile_update5 = PostTileUpdate(32, 16, 32, 4)
        self.tile_update6 = FinalTileUpdate(32, 3, 16, 2)

    def forward(self, left_fea_pyramid, right_fea_pyramid, init_tile_pyramid):
        if self.training:
            raise NotImplementedError(
                "codd_b200.TilePropagation is forward/inference only: the training-mode output pyramids "
                "(propagation.py:374-451) need autograd through the CUDA kernels (out of scope, DESIGN.md)")
        fl, fr, init = left_fea_pyramid, right_fea_pyramid, init_tile_pyramid
        t16 = self.tile_update0(fl[0], fr[0], init[0])
        t8 = self.tile_update1(fl[1], fr[1], init[1], t16[0])
        t4 = self.tile_update2(fl[2], fr[2], init[2], t8[0])
        t2 = self.tile_update3(fl[3], fr[3], init[3], t4[0])
        t1 = self.tile_update4(fl[4], fr[4], init[4], t2[0])
        r1 = self.tile_update4_1(fl[2], t1[0])
        r05 = self.tile_update5(fl[3], ops.plane_upsample(r1, 1.0, 2))
        r025 = self.tile_update6(fl[4], ops.plane_upsample(r05, 1.0, 2))
        return r025[:, 0:1, :, :]
