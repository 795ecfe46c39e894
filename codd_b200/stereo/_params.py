"""Packed-weight cache shared by the drop-in modules.

Parameters live in ordinary ``nn.Conv2d`` / ``nn.ConvTranspose2d`` containers so that
``state_dict`` names and shapes are the reference's; the CUDA kernels
read a re-laid-out copy ([tap][cin][cout]) that is rebuilt whenever the parameter changes
(``load_state_dict``, ``.to()``, in-place updates bump ``_version``).
"""
import os

import torch

from .. import ops

# 3x3 convolutions run on the tensor cores (tcgen05, 3xTF32) unless CODD_TC=0
USE_TC = os.environ.get("CODD_TC", "1") != "0"
# dilation-1 layers use the rolling-ring tcgen05 kernel (csrc/conv_tc_ring.cu) unless CODD_TC_RING=0
USE_RING = os.environ.get("CODD_TC_RING", "1") != "0"
# stacked 16-channel 3x3 pairs run as one fused ring launch (csrc/conv_tc_ring2.cu) unless CODD_TC_RING2=0
USE_RING2 = os.environ.get("CODD_TC_RING2", "1") != "0"


def run_conv(pw, conv, x, act, x2=None, residual=None, res_bcast=False, head=None):
    """One nn.Conv2d through the C ABI: tensor-core kernel when the layer is eligible, the fp32
    direct convolution otherwise.  ``head``: compute only the first ``head`` output channels."""
    cout = conv.out_channels if head is None else head
    k, st, pd, dl = conv.kernel_size, conv.stride, conv.padding, conv.dilation[0]
    cin = x.shape[1] + (0 if x2 is None else x2.shape[1])
    # rolling-ring kernel for everything but the tiny coarse-level layers (a strip segment needs ~16 rows per SM to
    # amortise its pipeline fill; measured cross-over at batch 8 between the 36x60 and 72x120 levels,
    # tools/conv_probe.py).  The choice depends on the per-sample size only: the two kernels round differently, and a
    # sample's result must not depend on the batch it is in.
    if cout == 1 and cin == 16 and x2 is None and k == (3, 3) and st == (1, 1) and pd == (1, 1) and dl == 1:
        # single-channel head (FinalTileUpdate disparity): memory-bound, its own shared-memory tile kernel
        wp, b = pw.conv_head(conv, 1)
        return ops.conv2d(x, wp, b, 1, k, st, pd, dl, act, residual=residual, res_bcast=res_bcast)
    if USE_TC and head is None and ops.tc4_eligible(x, cout, k, st, pd, dl, x2, residual):
        # conv_down first layer at the two finest levels (backbone.py:8-14): tcgen05 implicit GEMM over parity boxes
        ws, b = pw.conv_tc4(conv)
        return ops.conv4x4s2_tc(x, ws, b, cout, act)
    big = x.shape[2] * x.shape[3] >= 4096
    if (USE_TC and USE_RING and big and cout == 34 and cin == 32 and x2 is None and residual is None and dl == 1
            and act != ops.ACT_RELU_CH0 and ops.tc_eligible(cin, 32, k, st, pd, dl, x2)):
        # TileUpdate.lastconv (32 -> 34, propagation.py:190-199): the first 32 filters run on the tensor cores, the
        # last two as a direct convolution, both into one buffer with a 16-byte aligned pixel stride of 36 floats
        out = ops.empty_nhwc(x.shape[0], 34, x.shape[2], x.shape[3], x.device, 36)
        ws, b = pw.conv_ring(conv, 32)
        ops.conv3x3_tc_ring(x, ws, b, 32, act, out=out[:, :32])
        wp, b2 = pw.conv_range(conv, 32, 34)
        ops.conv2d(x, wp, b2, 2, k, st, pd, dl, act, out=out[:, 32:34])
        return out
    if (USE_TC and not big and cout == 34 and cin == 32 and x2 is None and residual is None and dl == 1
            and act != ops.ACT_RELU_CH0 and ops.tc_eligible(cin, 32, k, st, pd, dl, x2)):
        # the same layer at the coarse levels: halo-tile tensor-core kernel for the first 32 filters + the two-output head
        out = ops.empty_nhwc(x.shape[0], 34, x.shape[2], x.shape[3], x.device, 36)
        ws, b = pw.conv_tc(conv, 32)
        ops.conv3x3_tc(x, ws, b, 32, act, out=out[:, :32])
        wp, b2 = pw.conv_range(conv, 32, 34)
        ops.conv2d(x, wp, b2, 2, k, st, pd, dl, act, out=out[:, 32:34])
        return out
    if USE_TC and USE_RING and big and dl == 1 and ops.tc_eligible(cin, cout, k, st, pd, dl, x2):
        ws, b = pw.conv_ring(conv, head)
        return ops.conv3x3_tc_ring(x, ws, b, cout, act, residual=residual, res_bcast=res_bcast)
    if (USE_TC and USE_RING and big and dl == 3 and head is None and ops.tc_eligible(cin, cout, k, st, pd, dl, x2)
            and ops.ring_dil3_eligible(x, cout)):
        # dilated ResBlocks (propagation.py:258-280): the ring kernel on the three row phases of the image
        ws, b = pw.conv_ring(conv, None)
        return ops.conv3x3_tc_ring(x, ws, b, cout, act, residual=residual, res_bcast=res_bcast, dil=3)
    if USE_TC and ops.tc_eligible(cin, cout, k, st, pd, dl, x2):
        ws, b = pw.conv_tc(conv, head)
        return ops.conv3x3_tc(x, ws, b, cout, act, residual=residual, res_bcast=res_bcast, dil=dl)
    wp, b = pw.conv(conv) if head is None else pw.conv_head(conv, head)
    return ops.conv2d(x, wp, b, cout, k, st, pd, dl, act, x2=x2, residual=residual, res_bcast=res_bcast)


def run_conv_pair(pw, conv_a, act_a, conv_b, act_b, x, residual=None):
    """act_b(conv_b(act_a(conv_a(x))) [+ residual]) for two stacked 3x3 convolutions: one fused rolling-ring launch when
    both are 16 -> 16 (codd_conv3x3x2_tc_ring keeps the intermediate tensor on chip and is bit-identical to the two
    launches it replaces), two ``run_conv`` calls otherwise."""
    def plain3x3(c):
        return (c.kernel_size == (3, 3) and c.stride == (1, 1) and c.padding == (1, 1) and c.dilation == (1, 1)
                and c.in_channels == 16 and c.out_channels == 16)
    if (USE_TC and USE_RING and USE_RING2 and plain3x3(conv_a) and plain3x3(conv_b) and max(act_a, act_b) <= ops.ACT_RELU_CH0
            and ops.ring2_eligible(x, 16, 16, residual)):
        wa, ba = pw.conv_ring(conv_a, None)
        wb, bb = pw.conv_ring(conv_b, None)
        return ops.conv3x3x2_tc_ring(x, wa, ba, act_a, wb, bb, act_b, residual=residual)
    y = run_conv(pw, conv_a, x, act_a)
    return run_conv(pw, conv_b, y, act_b, residual=residual)


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

    def conv_range(self, conv, c0, c1):
        """Packed direct-conv weight / bias of output channels [c0, c1)."""
        w = conv.weight
        b = conv.bias
        key = (id(conv), "range", c0, c1)
        tag = (w.data_ptr(), w._version, w.device, None if b is None else (b.data_ptr(), b._version))
        hit = self._cache.get(key)
        if hit is None or hit[0] != tag:
            hit = (tag, ops.pack_conv_weight(w[c0:c1]), None if b is None else b.detach()[c0:c1].float().contiguous())
            self._cache[key] = hit
        return hit[1], hit[2]

    def conv_tc(self, conv, n=None):
        """hi/lo tf32 split of a 3x3 weight for the tensor-core kernel (first ``n`` filters if given)."""
        w = conv.weight
        b = conv.bias
        key = (id(conv), "tc", n)
        tag = (w.data_ptr(), w._version, w.device, None if b is None else (b.data_ptr(), b._version))
        hit = self._cache.get(key)
        if hit is None or hit[0] != tag:
            ww = w if n is None else w[:n]
            bb = None if b is None else (b.detach() if n is None else b.detach()[:n]).float().contiguous()
            hit = (tag, ops.pack_conv_weight_tc(ww), bb)
            self._cache[key] = hit
        return hit[1], hit[2]

    def conv_tc4(self, conv):
        """hi/lo tf32 split of a 4x4 stride-2 weight for codd_conv4x4s2_tc."""
        w = conv.weight
        b = conv.bias
        key = (id(conv), "tc4")
        tag = (w.data_ptr(), w._version, w.device, None if b is None else (b.data_ptr(), b._version))
        hit = self._cache.get(key)
        if hit is None or hit[0] != tag:
            hit = (tag, ops.pack_conv_weight_tc4(w), None if b is None else b.detach().float().contiguous())
            self._cache[key] = hit
        return hit[1], hit[2]

    def conv_ring(self, conv, n=None):
        """[pass][kx][6*NP][KC] layout of a 3x3 weight for the rolling-ring tensor-core kernel."""
        w = conv.weight
        b = conv.bias
        key = (id(conv), "ring", n)
        tag = (w.data_ptr(), w._version, w.device, None if b is None else (b.data_ptr(), b._version))
        hit = self._cache.get(key)
        if hit is None or hit[0] != tag:
            ww = w if n is None else w[:n]
            bb = None if b is None else (b.detach() if n is None else b.detach()[:n]).float().contiguous()
            hit = (tag, ops.pack_conv_weight_ring(ww), bb)
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
