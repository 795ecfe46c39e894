"""BasicEncoder — the RAFT3D feature network (model/motion/raft3d/blocks/extractor.py:9-55,124-199),
same parameter tree (conv1, layer{1,2,3}.{0,1}.{conv1,conv2,downsample.0}, conv2); instance-norm variant
(the only one RAFT3D builds, raft3d.py:149).  Forward = codd_conv2d_nhwc + codd_instance_norm_nhwc."""
import torch.nn as nn

from .. import ops
from ._net import NetWeights, conv


class ResidualBlock(nn.Module):
    def __init__(self, in_planes, planes, norm_fn="instance", stride=1):
        super().__init__()
        if norm_fn != "instance":
            raise NotImplementedError("codd_b200 BasicEncoder: instance norm only (what RAFT3D uses)")
        self.conv1 = nn.Conv2d(in_planes, planes, kernel_size=3, padding=1, stride=stride)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, padding=1)
        self.relu = nn.ReLU(inplace=True)
        self.norm1 = nn.InstanceNorm2d(planes)
        self.norm2 = nn.InstanceNorm2d(planes)
        if stride == 1:
            self.downsample = None
        else:
            self.norm3 = nn.InstanceNorm2d(planes)
            self.downsample = nn.Sequential(nn.Conv2d(in_planes, planes, kernel_size=1, stride=stride), self.norm3)

    def run(self, pw, x):
        y = ops.instance_norm(conv(pw, self.conv1, x), relu=True)
        y = conv(pw, self.conv2, y)
        if self.downsample is not None:
            x = ops.instance_norm(conv(pw, self.downsample[0], x), relu=False)
        return ops.instance_norm(y, relu=True, residual=x)       # relu(x + relu(norm2(conv2)))


class BasicEncoder(nn.Module):
    def __init__(self, output_dim=128, norm_fn="instance", dropout=0.0, depth_input=False):
        super().__init__()
        if norm_fn != "instance" or depth_input:
            raise NotImplementedError("codd_b200 BasicEncoder: instance norm, image input only")
        self.norm_fn = norm_fn
        self.norm1 = nn.InstanceNorm2d(64)
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3)
        self.relu1 = nn.ReLU(inplace=True)
        self.in_planes = 64
        self.layer1 = self._make_layer(64, stride=1)
        self.layer2 = self._make_layer(96, stride=2)
        self.layer3 = self._make_layer(128, stride=2)
        self.conv2 = nn.Conv2d(128, output_dim, kernel_size=1)
        self.dropout = None
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        self._pw = NetWeights()

    def _make_layer(self, dim, stride=1):
        layers = (ResidualBlock(self.in_planes, dim, self.norm_fn, stride=stride),
                  ResidualBlock(dim, dim, self.norm_fn, stride=1))
        self.in_planes = dim
        return nn.Sequential(*layers)

    def forward(self, x):
        pw = self._pw
        x = ops.instance_norm(conv(pw, self.conv1, ops.to_nhwc(x)), relu=True)
        for layer in (self.layer1, self.layer2, self.layer3):
            for blk in layer:
                x = blk.run(pw, x)
        return conv(pw, self.conv2, x)
