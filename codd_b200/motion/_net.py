"""Shared helpers of the RAFT3D network modules: convolutions through the C ABI with cached,
re-packed weights (BatchNorm folded in eval mode, several heads concatenated along Cout)."""
import torch

from .. import ops
from ..stereo._params import PackedWeights


def _tag(*tensors):
    return tuple(None if t is None else (t.data_ptr(), t._version, str(t.device)) for t in tensors)


class NetWeights(PackedWeights):
    def conv_bn(self, conv, bn):
        """conv (bias optional) followed by an eval-mode BatchNorm2d, folded into one packed weight + bias."""
        key = (id(conv), id(bn), "bn")
        tag = _tag(conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var)
        hit = self._cache.get(key)
        if hit is None or hit[0] != tag:
            w = conv.weight.detach().float()
            scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
            b = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
            if conv.bias is not None:
                b = b + conv.bias.detach().float() * scale
            hit = (tag, ops.pack_conv_weight(w * scale.view(-1, 1, 1, 1)), b.contiguous())
            self._cache[key] = hit
        return hit[1], hit[2]

    def conv_cat(self, convs):
        """Several convolutions of identical geometry reading the same input, as one wide convolution."""
        key = tuple(id(c) for c in convs) + ("cat",)
        tag = _tag(*[c.weight for c in convs], *[c.bias for c in convs])
        hit = self._cache.get(key)
        if hit is None or hit[0] != tag:
            w = torch.cat([c.weight.detach().float() for c in convs], 0)
            b = torch.cat([c.bias.detach().float() for c in convs], 0).contiguous()
            hit = (tag, ops.pack_conv_weight(w), b)
            self._cache[key] = hit
        return hit[1], hit[2]


def conv(pw, m, x, act=ops.ACT_NONE, bn=None, residual=None, out=None, x2=None):
    """One nn.Conv2d (optionally + eval BatchNorm) through codd_conv2d_nhwc."""
    wp, b = pw.conv_bn(m, bn) if bn is not None else pw.conv(m)
    return ops.conv2d(x, wp, b, m.out_channels, m.kernel_size, m.stride, m.padding, m.dilation[0], act, x2=x2,
                      residual=residual, out=out)
