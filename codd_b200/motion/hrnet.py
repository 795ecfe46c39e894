"""HRNet — the RAFT3D context backbone.  The reference builds mmseg's ``HRNet`` from the config in
configs/models/codd.py:48-73 (``builder_oss.build_backbone``, model/motion/raft3d/raft3d.py:155-158);
mmseg is not vendored in the reference tree, so this module restates its published structure
(stem 2x(3x3 s2, 64) -> stage1 Bottlenecks -> transitions / HRModules with fuse layers, eval-mode
BatchNorm) with the same parameter names (conv1, bn1, conv2, bn2, layer1.*, transition{1,2,3}.*,
stage{2,3,4}.*.branches.*, stage*.fuse_layers.*) so an mmseg checkpoint loads unchanged.
Parity unpinned.  Forward: BatchNorm folded into codd_conv2d_nhwc, fuse-layer
up-sampling by codd_resize_bilinear_nhwc (align_corners=False, accumulating)."""
import torch.nn as nn

from .. import ops
from ..registry import BACKBONES
from ._net import NetWeights, conv


def _bn(c):
    return nn.BatchNorm2d(c)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = _bn(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = _bn(planes)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def run(self, pw, x):
        idt = x if self.downsample is None else conv(pw, self.downsample[0], x, bn=self.downsample[1])
        y = conv(pw, self.conv1, x, ops.ACT_RELU, bn=self.bn1)
        return conv(pw, self.conv2, y, ops.ACT_RELU, bn=self.bn2, residual=idt)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = _bn(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = _bn(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = _bn(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def run(self, pw, x):
        idt = x if self.downsample is None else conv(pw, self.downsample[0], x, bn=self.downsample[1])
        y = conv(pw, self.conv1, x, ops.ACT_RELU, bn=self.bn1)
        y = conv(pw, self.conv2, y, ops.ACT_RELU, bn=self.bn2)
        return conv(pw, self.conv3, y, ops.ACT_RELU, bn=self.bn3, residual=idt)


BLOCKS = {"BASIC": BasicBlock, "BOTTLENECK": Bottleneck}


def make_layer(block, inplanes, planes, blocks, stride=1):
    downsample = None
    if stride != 1 or inplanes != planes * block.expansion:
        downsample = nn.Sequential(nn.Conv2d(inplanes, planes * block.expansion, 1, stride=stride, bias=False),
                                   _bn(planes * block.expansion))
    layers = [block(inplanes, planes, stride, downsample)]
    inplanes = planes * block.expansion
    for _ in range(1, blocks):
        layers.append(block(inplanes, planes))
    return nn.Sequential(*layers)


class HRModule(nn.Module):
    def __init__(self, num_branches, block, num_blocks, in_channels, num_channels, multiscale_output=True):
        super().__init__()
        self.in_channels = list(in_channels)
        self.num_branches = num_branches
        self.multiscale_output = multiscale_output
        branches = []
        for i in range(num_branches):
            branches.append(make_layer(block, self.in_channels[i], num_channels[i], num_blocks[i]))
            self.in_channels[i] = num_channels[i] * block.expansion
        self.branches = nn.ModuleList(branches)
        self.fuse_layers = self._make_fuse_layers()
        self.relu = nn.ReLU(inplace=False)

    def _make_fuse_layers(self):
        if self.num_branches == 1:
            return None
        c = self.in_channels
        fuse_layers = []
        for i in range(self.num_branches if self.multiscale_output else 1):
            row = []
            for j in range(self.num_branches):
                if j > i:
                    row.append(nn.Sequential(nn.Conv2d(c[j], c[i], 1, bias=False), _bn(c[i]),
                                             nn.Upsample(scale_factor=2 ** (j - i), mode="bilinear", align_corners=False)))
                elif j == i:
                    row.append(None)
                else:
                    downs = []
                    for k in range(i - j):
                        if k == i - j - 1:
                            downs.append(nn.Sequential(nn.Conv2d(c[j], c[i], 3, stride=2, padding=1, bias=False), _bn(c[i])))
                        else:
                            downs.append(nn.Sequential(nn.Conv2d(c[j], c[j], 3, stride=2, padding=1, bias=False), _bn(c[j]),
                                                       nn.ReLU(inplace=False)))
                    row.append(nn.Sequential(*downs))
            fuse_layers.append(nn.ModuleList(row))
        return nn.ModuleList(fuse_layers)

    def run(self, pw, xs):
        xs = list(xs)
        for i in range(self.num_branches):
            for blk in self.branches[i]:
                xs[i] = blk.run(pw, xs[i])
        if self.num_branches == 1:
            return xs
        outs = []
        nb = self.num_branches
        for i, row in enumerate(self.fuse_layers):
            # y = sum_j f_ij(x_j), accumulated in the reference's order j = 0..nb-1; ReLU on the last term
            y = None
            for j in range(nb):
                last = j == nb - 1
                if j == i:
                    if y is None:
                        y = xs[j]
                        if last:
                            y = ops.eltwise(ops.EW_ACT, y, act=ops.ACT_RELU)
                    else:
                        y = ops.eltwise(ops.EW_ADD_ACT, y, xs[j], act=ops.ACT_RELU if last else ops.ACT_NONE)
                elif j > i:
                    t = conv(pw, row[j][0], xs[j], bn=row[j][1])
                    y = ops.resize_bilinear(t, xs[i].shape[2:], False, base=y, relu=last)
                else:
                    t = xs[j]
                    downs = row[j]
                    for k, d in enumerate(downs):
                        if k == len(downs) - 1:
                            t = conv(pw, d[0], t, ops.ACT_RELU if last else ops.ACT_NONE, bn=d[1], residual=y)
                        else:
                            t = conv(pw, d[0], t, ops.ACT_RELU, bn=d[1])
                    y = t
            outs.append(y)
        return outs


@BACKBONES.register_module(force=True)
class HRNet(nn.Module):
    def __init__(self, extra, in_channels=3, conv_cfg=None, norm_cfg=None, norm_eval=False, with_cp=False,
                 frozen_stages=-1, zero_init_residual=False, multiscale_output=True, pretrained=None, init_cfg=None):
        super().__init__()
        self.extra = extra
        self.norm_eval = norm_eval
        self.conv1 = nn.Conv2d(in_channels, 64, 3, stride=2, padding=1, bias=False)
        self.bn1 = _bn(64)
        self.conv2 = nn.Conv2d(64, 64, 3, stride=2, padding=1, bias=False)
        self.bn2 = _bn(64)
        self.relu = nn.ReLU(inplace=True)

        s1 = extra["stage1"]
        block = BLOCKS[s1["block"]]
        stage1_out = s1["num_channels"][0] * block.expansion
        self.layer1 = make_layer(block, 64, s1["num_channels"][0], s1["num_blocks"][0])
        pre = [stage1_out]
        for idx in (2, 3, 4):
            cfg = extra[f"stage{idx}"]
            block = BLOCKS[cfg["block"]]
            chans = [c * block.expansion for c in cfg["num_channels"]]
            setattr(self, f"transition{idx - 1}", self._make_transition_layer(pre, chans))
            stage, pre = self._make_stage(cfg, chans, multiscale_output if idx == 4 else True)
            setattr(self, f"stage{idx}", stage)
        self._pw = NetWeights()

    @staticmethod
    def _make_transition_layer(pre, cur):
        layers = []
        for i in range(len(cur)):
            if i < len(pre):
                if cur[i] != pre[i]:
                    layers.append(nn.Sequential(nn.Conv2d(pre[i], cur[i], 3, padding=1, bias=False), _bn(cur[i]),
                                                nn.ReLU(inplace=True)))
                else:
                    layers.append(None)
            else:
                downs = []
                for j in range(i + 1 - len(pre)):
                    cin = pre[-1]
                    cout = cur[i] if j == i - len(pre) else cin
                    downs.append(nn.Sequential(nn.Conv2d(cin, cout, 3, stride=2, padding=1, bias=False), _bn(cout),
                                               nn.ReLU(inplace=True)))
                layers.append(nn.Sequential(*downs))
        return nn.ModuleList(layers)

    @staticmethod
    def _make_stage(cfg, in_channels, multiscale_output=True):
        block = BLOCKS[cfg["block"]]
        mods = []
        for i in range(cfg["num_modules"]):
            ms = multiscale_output or i != cfg["num_modules"] - 1
            m = HRModule(cfg["num_branches"], block, cfg["num_blocks"], in_channels, cfg["num_channels"], ms)
            in_channels = m.in_channels
            mods.append(m)
        return nn.Sequential(*mods), in_channels

    def init_weights(self):
        pass

    def train(self, mode=True):
        super().train(mode)
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, nn.BatchNorm2d):
                    m.eval()
        return self

    def _transition(self, pw, layer, x):
        if isinstance(layer[0], nn.Conv2d):
            return conv(pw, layer[0], x, ops.ACT_RELU, bn=layer[1])
        for d in layer:
            x = conv(pw, d[0], x, ops.ACT_RELU, bn=d[1])
        return x

    def forward(self, x):
        if self.training and not self.norm_eval:
            raise NotImplementedError("codd_b200 HRNet: eval-mode BatchNorm only (inference path)")
        pw = self._pw
        x = conv(pw, self.conv1, ops.to_nhwc(x), ops.ACT_RELU, bn=self.bn1)
        x = conv(pw, self.conv2, x, ops.ACT_RELU, bn=self.bn2)
        for blk in self.layer1:
            x = blk.run(pw, x)
        ys = [x]
        for idx in (2, 3, 4):
            trans = getattr(self, f"transition{idx - 1}")
            xs = []
            for i, t in enumerate(trans):
                if t is not None:
                    xs.append(self._transition(pw, t, ys[-1] if idx > 2 else x))
                else:
                    xs.append(ys[i])
            for m in getattr(self, f"stage{idx}"):
                xs = m.run(pw, xs)
            ys = xs
        return ys
