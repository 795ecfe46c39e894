"""Packed-weight cache shared by the drop-in modules.

Parameters live in ordinary ``nn.Conv2d`` / ``nn.ConvTranspose2d`` containers so that
``state_dict`` names and shapes are the reference's; the CUDA kernels
read a re-laid-out copy ([tap][cin][cout]) that is rebuilt whenever the parameter changes
(``load_state_dict``, ``.to()``, in-place updates bump ``_version``).
"""
import torch

from .. import ops


class PackedWeights:
    def __init__(self):
        self._cache = {}

    def _get(self, conv, packer):
        w = conv.weight
        key = id(conv)
        b = conv.bias
        tag = (w.data_ptr(), w._version, w.device, None if b is None else (b.data_ptr(), b._version))
        hit = self._cache.get(key)
        if hit is None or hit[0] != tag:
            hit = (tag, packer(w), None if conv.bias is None else conv.bias.detach().float().contiguous())
            self._cache[key] = hit
        return hit[1], hit[2]

    def conv(self, conv):
        return self._get(conv, ops.pack_conv_weight)

    def conv_head(self, conv, n):
        """Packed weight / bias of the first ``n`` output channels only (eval-time heads that
        compute more channels than the forward pass consumes)."""
        w = conv.weight
        b = conv.bias
        key = (id(conv), n)
        tag = (w.data_ptr(), w._version, w.device, None if b is None else (b.data_ptr(), b._version))
        hit = self._cache.get(key)
        if hit is None or hit[0] != tag:
            hit = (tag, ops.pack_conv_weight(w[:n]), None if b is None else b.detach()[:n].float().contiguous())
            self._cache[key] = hit
        return hit[1], hit[2]

    def deconv(self, conv):
        return self._get(conv, ops.pack_deconv_weight)

    def raw(self, conv):
        """torch-layout weight flattened (1x1 convs read directly by a fused kernel)."""
        return self._get(conv, lambda w: w.detach().float().contiguous())


def kpad(conv):
    """(kernel, stride, padding, dilation) of an nn.Conv2d as plain ints / tuples."""
    return conv.kernel_size, conv.stride, conv.padding, conv.dilation[0]
