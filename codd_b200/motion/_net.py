"""Shared helpers of the RAFT3D network modules: convolutions through the C ABI with cached,
re-packed weights (BatchNorm folded in eval mode, several heads concatenated along Cout)."""
import os

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

    def conv_gemm_bn(self, conv, bn):
        """conv + eval-mode BatchNorm2d folded, packed for codd_gemm_tc (the wide HRNet layers)."""
        key = (id(conv), id(bn), "gemm_bn")
        tag = _tag(conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var)
        hit = self._cache.get(key)
        if hit is None or hit[0] != tag:
            w = conv.weight.detach().float()
            scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
            b = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
            if conv.bias is not None:
                b = b + conv.bias.detach().float() * scale
            hit = (tag, ops.pack_conv_weight_gemm(w * scale.view(-1, 1, 1, 1)), b.contiguous())
            self._cache[key] = hit
        return hit[1], hit[2]

    def conv_gemm_cat(self, convs):
        """[Cout][taps*Cin] tf32 hi / lo halves of one or several (concatenated) convolutions for codd_gemm_tc."""
        key = tuple(id(c) for c in convs) + ("gemm",)
        tag = _tag(*[c.weight for c in convs], *[c.bias for c in convs])
        hit = self._cache.get(key)
        if hit is None or hit[0] != tag:
            w = torch.cat([c.weight.detach().float() for c in convs], 0)
            b = torch.cat([c.bias.detach().float() for c in convs], 0).contiguous()
            hit = (tag, ops.pack_conv_weight_gemm(w), b)
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


USE_GEMM = os.environ.get("CODD_GEMM", "1") != "0"   # wide stride-1 layers on the tcgen05 GEMM (csrc/gemm_tc.cu)


def conv_cat(pw, convs, x, act=ops.ACT_NONE, residual=None):
    """Several same-geometry convolutions of one input as ONE wide convolution (Cout concatenated)."""
    m = convs[0]
    cout = sum(c.out_channels for c in convs)
    n, cin, h, w = x.shape
    if USE_GEMM and ops.gemm_eligible(n, h, w, cin, cout, m.kernel_size, m.stride, None):
        wg, b = pw.conv_gemm_cat(convs)
        return ops.conv2d_gemm(x, wg, b, cout, m.kernel_size, m.padding, m.dilation[0], act, residual=residual)
    wp, b = pw.conv_cat(convs)
    return ops.conv2d(x, wp, b, cout, m.kernel_size, m.stride, m.padding, m.dilation[0], act, residual=residual)


def conv(pw, m, x, act=ops.ACT_NONE, bn=None, residual=None, out=None, x2=None):
    """One nn.Conv2d (optionally + eval BatchNorm): tcgen05 GEMM for the wide stride-1 layers of the update block,
    codd_conv2d_nhwc otherwise."""
    n, cin, h, w = x.shape
    if (USE_GEMM and bn is None and out is None and m.bias is not None
            and ops.gemm_eligible(n, h, w, cin, m.out_channels, m.kernel_size, m.stride, x2)):
        wg, b = pw.conv_gemm_cat([m])
        return ops.conv2d_gemm(x, wg, b, m.out_channels, m.kernel_size, m.padding, m.dilation[0], act, residual=residual)
    if (USE_GEMM and bn is not None and out is None and ops.gemm_eligible(n, h, w, cin, m.out_channels, m.kernel_size, m.stride, x2)
            and 2 * m.padding[0] == m.dilation[0] * (m.kernel_size[0] - 1) and 2 * m.padding[1] == m.dilation[1] * (m.kernel_size[1] - 1)):
        # the wide stride-1 layers of HRNet (64 / 72 / 144 channels at 1/4 .. 1/32 resolution): BatchNorm folded, tcgen05 GEMM
        wg, b = pw.conv_gemm_bn(m, bn)
        return ops.conv2d_gemm(x, wg, b, m.out_channels, m.kernel_size, m.padding, m.dilation[0], act, residual=residual)
    wp, b = pw.conv_bn(m, bn) if bn is not None else pw.conv(m)
    return ops.conv2d(x, wp, b, m.out_channels, m.kernel_size, m.stride, m.padding, m.dilation[0], act, x2=x2,
                      residual=residual, out=out)
