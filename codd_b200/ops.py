"""Tensor-level wrappers over the C ABI.

Every function takes / returns torch CUDA tensors whose *logical* shape is the reference's
NCHW but whose memory is NHWC (``torch.channels_last``, possibly a channel slice of a wider
buffer).  Torch is used for allocation, streams and nothing else: all arithmetic happens in
``libcodd_b200.so``.  Non-CUDA inputs raise — there is no CPU path.
"""
import ctypes
import os

import torch

from . import lib as _lib
from .lib import ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_RELU_CH0, ACT_SIGMOID, ACT_TANH, ConvDesc  # noqa: F401

LAUNCHES = [0]  # kernels launched through this module (bench.py reports it as gpu_launches)
PROFILE = None   # when a list: (tag, algorithmic bytes, start event, end event) per launch


def _run(tag, nbytes, call):
    """Launch bookkeeping: counts the launch and, in a profiling pass, brackets it with CUDA
    events on the launching stream (torch's current stream)."""
    LAUNCHES[0] += 1
    if PROFILE is None:
        return call()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    rc = call()
    e.record()
    PROFILE.append((tag, nbytes, s, e))
    return rc


class profile:
    """``with ops.profile() as p:`` ... ``p.summary()`` -> per-kernel time / algorithmic bytes."""

    def __enter__(self):
        global PROFILE
        self.records = PROFILE = []
        return self

    def __exit__(self, *exc):
        global PROFILE
        PROFILE = None
        return False

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for tag, nbytes, s, e in self.records:
            a = agg.setdefault(tag, dict(kernel=tag, launches=0, ms=0.0, bytes=0))
            a["launches"] += 1
            a["ms"] += s.elapsed_time(e)
            a["bytes"] += nbytes
        return sorted(agg.values(), key=lambda a: -a["ms"])


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and (not t.is_cuda or t.dtype != torch.float32):
            raise _lib.CoddError("codd_b200 ops need float32 CUDA tensors (no CPU fallback)")


def empty_nhwc(n, c, h, w, device, ld=None):
    """Logical [n,c,h,w] tensor backed by NHWC memory with pixel stride ``ld`` (default c)."""
    ld = c if ld is None else ld
    buf = torch.empty((n, h, w, ld), device=device, dtype=torch.float32)
    t = buf.permute(0, 3, 1, 2)
    return t if ld == c else t[:, :c]


def ld_of(t):
    """Pixel stride of an NHWC-backed logical-NCHW tensor (validates the layout)."""
    n, c, h, w = t.shape
    s = t.stride()
    if c > 1 and s[1] != 1:
        raise _lib.CoddError(f"tensor is not NHWC-backed (strides {s}); call ops.to_nhwc first")
    if w > 1:
        ld = s[3]
    elif h > 1:
        ld = s[2]
    elif n > 1:
        ld = s[0]
    else:
        ld = c
    ok = (w == 1 or s[3] == ld) and (h == 1 or s[2] == w * ld) and (n == 1 or s[0] == h * w * ld) and ld >= c
    if not ok:
        raise _lib.CoddError(f"tensor is not NHWC-backed (shape {tuple(t.shape)}, strides {s})")
    return ld


def to_nhwc(t):
    """Accept any float32 CUDA NCHW tensor; returns an NHWC-backed equivalent (kernel copy if needed)."""
    _require_cuda(t)
    try:
        ld_of(t)
        return t
    except _lib.CoddError:
        pass
    t = t.contiguous()
    n, c, h, w = t.shape
    out = empty_nhwc(n, c, h, w, t.device)
    rc = _run("nchw_to_nhwc", 8 * t.numel(),
              lambda: _lib.load().codd_nchw_to_nhwc(t.data_ptr(), n, c, h, w, out.data_ptr(), c, _stream()))
    _lib.check(rc, "codd_nchw_to_nhwc")
    return out


def copy_to_nhwc(src, out):
    """Copy any float32 CUDA logical-NCHW tensor into an NHWC-backed destination (e.g. a channel slice)."""
    _require_cuda(src, out)
    try:
        ld_of(src)
    except _lib.CoddError:
        src = src.contiguous()
        n, c, h, w = src.shape
        rc = _run("nchw_to_nhwc", 8 * src.numel(), lambda: _lib.load().codd_nchw_to_nhwc(
            src.data_ptr(), n, c, h, w, out.data_ptr(), ld_of(out), _stream()))
        _lib.check(rc, "codd_nchw_to_nhwc")
        return out
    return eltwise(0, src, out=out)


def to_nchw(t):
    """NHWC-backed logical-NCHW tensor -> plain contiguous NCHW tensor (planar input: returned as is)."""
    _require_cuda(t)
    if _is_planar(t):
        return t
    n, c, h, w = t.shape
    ld = ld_of(t)
    out = torch.empty((n, c, h, w), device=t.device, dtype=torch.float32)
    rc = _run("nhwc_to_nchw", 8 * out.numel(),
              lambda: _lib.load().codd_nhwc_to_nchw(t.data_ptr(), ld, n, h, w, c, out.data_ptr(), _stream()))
    _lib.check(rc, "codd_nhwc_to_nchw")
    return out


def pack_conv_weight(w):
    """torch [Cout,Cin,KH,KW] -> [KH*KW][Cin][Cout] (what codd_conv2d_nhwc reads)."""
    return w.detach().permute(2, 3, 1, 0).contiguous().float()


def pack_deconv_weight(w):
    """torch ConvTranspose2d [Cin,Cout,2,2] -> [4][Cin][Cout]."""
    return w.detach().permute(2, 3, 0, 1).contiguous().float()


def _tf32_round(t):
    """Round-to-nearest (ties away) to the 10-bit tf32 mantissa, kept in an fp32 container."""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def pack_conv_weight_tc(w):
    """torch [Cout,Cin,3,3] -> [2][9][NP][KC] hi/lo split for codd_conv3x3_tc (3xTF32)."""
    cout, cin, kh, kw = w.shape
    assert kh == 3 and kw == 3
    kc = 16 if cin <= 16 else 32
    npad = 16 if cout <= 16 else 32
    wt = torch.zeros((9, npad, kc), dtype=torch.float32, device=w.device)
    wt[:, :cout, :cin] = w.detach().float().permute(2, 3, 0, 1).reshape(9, cout, cin)
    hi = _tf32_round(wt)
    lo = _tf32_round(wt - hi)
    return torch.stack([hi, lo]).contiguous()


def pack_conv_weight_tc4(w):
    """torch [Cout,Cin,4,4] -> fp16 [16 taps][2 NP rows][KC] for codd_conv4x4s2_tc / codd_tile_features_tc (returned
    viewed as float32): per tap NP rows of w_hi = fp16(w) followed by NP rows of fp16(2^10 (w - w_hi)); tap = ky*4 + kx,
    KC = 16 (Cin = 16) or 32 (Cin = 24, 32: zero-padded), NP = 16 | 32.  Same three-product scheme as
    pack_conv_weight_ring."""
    cout, cin, kh, kw = w.shape
    assert kh == 4 and kw == 4 and cin in (16, 24, 32) and cout <= 32
    npad = 16 if cout <= 16 else 32
    kc = 16 if cin <= 16 else 32
    wt = torch.zeros((16, npad, kc), dtype=torch.float32, device=w.device)
    wt[:, :cout, :cin] = w.detach().float().permute(2, 3, 0, 1).reshape(16, cout, cin)
    wt = wt.clamp(-65504.0, 65504.0)
    hi = wt.half()
    lo = ((wt - hi.float()) * 1024.0).half()
    return torch.cat([hi, lo], dim=1).contiguous().view(torch.float32)     # [16][2 NP][KC / 2] as float32


def tc4_eligible(x, cout, k, stride, pad, dil, x2, residual):
    """4x4 / stride 2 / pad 1, Cin in {16, 24, 32}, Cout <= 32, even sizes, large enough to fill the persistent grid
    (a per-sample criterion: a sample's result must not depend on the batch it is in)."""
    n, cin, h, w = x.shape
    return (x2 is None and residual is None and tuple(k) == (4, 4) and tuple(stride) == (2, 2) and tuple(pad) == (1, 1)
            and dil == 1 and cin in (16, 24, 32) and cout <= 32 and h % 2 == 0 and w % 2 == 0 and h * w >= 4096)


def conv4x4s2_tc(x, wsplit, bias, cout, act=ACT_NONE):
    """4x4 s2 p1 conv on the tensor cores (three fp16 products, fp32 accumulation).  ``wsplit`` from pack_conv_weight_tc4."""
    _require_cuda(x, wsplit, bias)
    n, cin, h, w = x.shape
    out = empty_nhwc(n, cout, h // 2, w // 2, x.device)
    nbytes = 4 * (n * h * w * cin + n * (h // 2) * (w // 2) * cout + wsplit.numel())
    rc = _run(f"conv4x4s2tc_cin{cin}_cout{cout}", nbytes, lambda: _lib.load().codd_conv4x4s2_tc(
        x.data_ptr(), ld_of(x), cin, n, h, w, wsplit.data_ptr(), None if bias is None else bias.data_ptr(), cout, act,
        out.data_ptr(), ld_of(out), _stream()))
    _lib.check(rc, f"codd_conv4x4s2_tc(cin={cin}, cout={cout})")
    return out


def pack_conv_weight_ring(w):
    """torch [Cout,Cin,3,3] -> flat buffer for codd_conv3x3_tc_ring (fp16 data, returned viewed as float32):
    pass A  [3 kx][6*NP rows][KC] — rows per ky = [w_hi (NP) | 2^10 * w_lo (NP)],  w_hi = fp16(w), w_lo = w - w_hi,
    pass B  [3 kx][6*NP rows][KC] — rows per ky = [0 (NP) | w_hi (NP)].
    With x split the same way by the kernel, the lo half of a TMEM slot accumulates 2^10 * (x_hi*w_lo + x_lo*w_hi) and
    the epilogue scales it back: three fp16 products with fp32 accumulation = fp32-class accuracy (cf. 3xTF32)."""
    cout, cin, kh, kw = w.shape
    assert kh == 3 and kw == 3
    kc = 16 if cin <= 16 else 32
    npad = 16 if cout <= 16 else 32
    wt = torch.zeros((3, 3, npad, kc), dtype=torch.float32, device=w.device)            # [ky][kx][cout][cin]
    wt[:, :, :cout, :cin] = w.detach().float().permute(2, 3, 0, 1)
    wt = wt.clamp(-65504.0, 65504.0)
    hi = wt.half()
    lo = ((wt - hi.float()) * 1024.0).half()
    pa = torch.zeros((3, 3, 2, npad, kc), dtype=torch.float16, device=w.device)         # [kx][ky][hi|lo][cout][cin]
    pa[:, :, 0] = hi.permute(1, 0, 2, 3)
    pa[:, :, 1] = lo.permute(1, 0, 2, 3)
    pb = torch.zeros((3, 3, 2, npad, kc), dtype=torch.float16, device=w.device)
    pb[:, :, 1] = hi.permute(1, 0, 2, 3)
    return torch.cat([pa.reshape(-1), pb.reshape(-1)]).contiguous().view(torch.float32)


def ring_dil3_eligible(x, cout):
    """Dilation-3 layers the rolling-ring kernel takes (32 -> 32 channels, image height a multiple of 3)."""
    return x.shape[1] == 32 and cout == 32 and x.shape[2] % 3 == 0


def conv3x3_tc_ring(x, wring, bias, cout, act=ACT_NONE, residual=None, res_bcast=False, out=None, dil=1):
    """3x3 s1 conv (pad = dil) on the tensor cores, rolling-ring kernel (3 fp16 products, fp32 accumulation).  ``wring``
    from pack_conv_weight_ring; ``out``: optional NHWC-backed destination (e.g. a channel slice of a wider buffer);
    ``dil`` = 3 needs ring_dil3_eligible."""
    _require_cuda(x, wring, bias, residual, out)
    n, cin, h, w = x.shape
    if out is None:
        out = empty_nhwc(n, cout, h, w, x.device)
    nbytes = 4 * (n * h * w * (cin + cout) + wring.numel()
                  + (0 if residual is None else n * h * w * (1 if res_bcast else cout)))
    tag = f"conv3x3ring_cin{cin}_cout{cout}" + ("" if dil == 1 else f"_d{dil}")
    rc = _run(tag, nbytes, lambda: _lib.load().codd_conv3x3_tc_ring_dil(
        x.data_ptr(), ld_of(x), cin, n, h, w, wring.data_ptr(), None if bias is None else bias.data_ptr(),
        None if residual is None else residual.data_ptr(), 0 if residual is None else ld_of(residual),
        1 if res_bcast else 0, cout, act, out.data_ptr(), ld_of(out), dil, _stream()))
    _lib.check(rc, f"codd_conv3x3_tc_ring_dil(cin={cin}, cout={cout}, dil={dil})")
    return out


def ring2_eligible(x, cmid, cout, residual=None):
    """Two stacked 16-channel 3x3 convolutions in one launch (codd_conv3x3x2_tc_ring): the same per-sample size rule as
    the ring kernel, 32-byte aligned NHWC rows."""
    return (x.shape[1] == 16 and cmid == 16 and cout == 16 and x.shape[2] * x.shape[3] >= 4096
            and (residual is None or (ld_of(residual) % 8 == 0 and residual.data_ptr() % 32 == 0)))


def conv3x3x2_tc_ring(x, wring_a, bias_a, act_a, wring_b, bias_b, act_b, residual=None):
    """act_b(conv_b(act_a(conv_a(x))) [+ residual]) for 16 -> 16 -> 16 channels, the intermediate kept on chip."""
    _require_cuda(x, wring_a, bias_a, wring_b, bias_b, residual)
    n, cin, h, w = x.shape
    out = empty_nhwc(n, 16, h, w, x.device)
    nbytes = 4 * (n * h * w * (32 + (0 if residual is None else 16)) + wring_a.numel() + wring_b.numel())
    rc = _run("conv3x3x2ring_c16", nbytes, lambda: _lib.load().codd_conv3x3x2_tc_ring(
        x.data_ptr(), ld_of(x), n, h, w, wring_a.data_ptr(), None if bias_a is None else bias_a.data_ptr(), act_a,
        wring_b.data_ptr(), None if bias_b is None else bias_b.data_ptr(),
        None if residual is None else residual.data_ptr(), 0 if residual is None else ld_of(residual), act_b,
        out.data_ptr(), ld_of(out), _stream()))
    _lib.check(rc, "codd_conv3x3x2_tc_ring")
    return out


def pack_conv_weight_gemm(w):
    """torch [Cout,Cin,KH,KW] -> (B_hi, B_lo), each [Cout][KH*KW*Cin] (k = tap*Cin + ch), tf32 halves for codd_gemm_tc."""
    cout = w.shape[0]
    b = w.detach().float().permute(0, 2, 3, 1).reshape(cout, -1).contiguous()
    hi = _tf32_round(b)
    return hi.contiguous(), _tf32_round(b - hi).contiguous()


GEMM_3XTF32 = True    # False: single-pass TF32 (what cuDNN does by default for the reference on GPUs)


def gemm_eligible(n, h, w, cin, cout, k, stride, x2):
    kh, kw = (k, k) if isinstance(k, int) else k
    sh, sw = (stride, stride) if isinstance(stride, int) else stride
    m, kk = n * h * w, kh * kw * cin
    return (x2 is None and sh == 1 and sw == 1 and cin % 4 == 0 and kk >= 128 and cout >= 16 and m >= 1024
            and m * kk * 8 <= (3 << 29))


def conv2d_gemm(x, wg, bias, cout, k, pad=(0, 0), dil=1, act=ACT_NONE, residual=None, out=None):
    """Stride-1 convolution as im2col + tcgen05 GEMM (3xTF32 unless ops.GEMM_3XTF32 is False).  ``wg`` from
    pack_conv_weight_gemm; out = act(conv(x) + bias + residual)."""
    _require_cuda(x, wg[0], wg[1], bias, residual, out)
    n, cin, h, w = x.shape
    kh, kw = (k, k) if isinstance(k, int) else k
    ph, pw = (pad, pad) if isinstance(pad, int) else pad
    if 2 * ph != dil * (kh - 1) or 2 * pw != dil * (kw - 1):
        raise _lib.CoddError("conv2d_gemm: 'same' padding only")
    m, kk = n * h * w, kh * kw * cin
    if out is None:
        out = empty_nhwc(n, cout, h, w, x.device)
    a = torch.empty((m, kk), device=x.device, dtype=torch.float32)
    a_lo = torch.empty_like(a) if GEMM_3XTF32 else None
    rc = _run(f"im2col_k{kh}x{kw}_cin{cin}", 4 * (x.numel() + (2 if GEMM_3XTF32 else 1) * m * kk), lambda: _lib.load().codd_im2col_split(
        x.data_ptr(), ld_of(x), n, h, w, cin, kh, kw, ph, pw, dil, a.data_ptr(), None if a_lo is None else a_lo.data_ptr(), kk,
        _stream()))
    _lib.check(rc, "codd_im2col_split")
    nseg = 3 if GEMM_3XTF32 else 1
    nbytes = 4 * (nseg * m * kk + cout * kk + m * cout * (1 if residual is None else 2))
    rc = _run(f"gemm_tc_k{kk}_n{cout}", nbytes, lambda: _lib.load().codd_gemm_tc(
        a.data_ptr(), None if a_lo is None else a_lo.data_ptr(), kk, wg[0].data_ptr(),
        wg[1].data_ptr() if GEMM_3XTF32 else None, kk, m, cout, kk, None if bias is None else bias.data_ptr(),
        None if residual is None else residual.data_ptr(), 0 if residual is None else ld_of(residual), act, out.data_ptr(),
        ld_of(out), _stream()))
    _lib.check(rc, f"codd_gemm_tc(m={m}, n={cout}, k={kk})")
    return out


def tc_eligible(cin, cout, k, stride, pad, dil, x2):
    kh, kw = (k, k) if isinstance(k, int) else k
    sh, sw = (stride, stride) if isinstance(stride, int) else stride
    ph, pw = (pad, pad) if isinstance(pad, int) else pad
    if not (x2 is None and kh == 3 and kw == 3 and sh == 1 and sw == 1 and ph == dil and pw == dil):
        return False
    if dil == 3:
        return cin == 32 and cout == 32
    return dil == 1 and cin in (16, 24, 32) and cout <= 32 and not (cin <= 16 and cout > 16)


def conv3x3_tc(x, wsplit, bias, cout, act=ACT_NONE, residual=None, res_bcast=False, flags=0, dil=1, out=None):
    """3x3 s1 conv (pad = dil) on the tensor cores (3xTF32).  ``wsplit`` from pack_conv_weight_tc; ``out``: optional
    NHWC-backed destination (e.g. a channel slice of a wider buffer)."""
    _require_cuda(x, wsplit, bias, residual, out)
    n, cin, h, w = x.shape
    if out is None:
        out = empty_nhwc(n, cout, h, w, x.device)
    nbytes = 4 * (n * h * w * (cin + cout) + wsplit.numel() // 2
                  + (0 if residual is None else n * h * w * (1 if res_bcast else cout)))
    tag = f"conv3x3tc_cin{cin}_cout{cout}" + ("" if dil == 1 else f"_d{dil}")
    rc = _run(tag, nbytes, lambda: _lib.load().codd_conv3x3_tc_dil(
        x.data_ptr(), ld_of(x), cin, n, h, w, wsplit.data_ptr(), None if bias is None else bias.data_ptr(),
        None if residual is None else residual.data_ptr(), 0 if residual is None else ld_of(residual),
        1 if res_bcast else 0, cout, act, out.data_ptr(), ld_of(out), dil, flags, _stream()))
    _lib.check(rc, f"codd_conv3x3_tc(cin={cin}, cout={cout}, dil={dil})")
    return out


def conv2d(x, wp, bias, cout, k, stride=(1, 1), pad=(0, 0), dil=1, act=ACT_NONE, x2=None, residual=None,
           res_bcast=False, out=None, out_hw=None, ld_out=None, res_after_act=False):
    """act(conv(cat(x, x2)) + bias + residual).  ``wp`` is a packed weight."""
    _require_cuda(x, x2, wp, bias, residual, out)
    n, c0, h, w = x.shape
    kh, kw = (k, k) if isinstance(k, int) else k
    sh, sw = (stride, stride) if isinstance(stride, int) else stride
    ph, pw = (pad, pad) if isinstance(pad, int) else pad
    if out_hw is None:
        ho = (h + 2 * ph - dil * (kh - 1) - 1) // sh + 1
        wo = (w + 2 * pw - dil * (kw - 1) - 1) // sw + 1
    else:
        ho, wo = out_hw
    c1 = 0 if x2 is None else x2.shape[1]
    if wp.numel() != kh * kw * (c0 + c1) * cout:
        raise _lib.CoddError(f"packed weight has {wp.numel()} elements, expected {kh*kw*(c0+c1)*cout}")
    if out is None:
        out = empty_nhwc(n, cout, ho, wo, x.device, ld_out)
    d = ConvDesc(n=n, h=h, w=w, c0=c0, ld0=ld_of(x), c1=c1, ld1=0 if x2 is None else ld_of(x2), cout=cout,
                 ldo=ld_of(out), kh=kh, kw=kw, sh=sh, sw=sw, ph=ph, pw=pw, dil=dil, ho=ho, wo=wo, act=act,
                 ldr=0 if residual is None else ld_of(residual), res_bcast=1 if res_bcast else 0,
                 res_after_act=1 if res_after_act else 0)
    nbytes = 4 * (n * h * w * (c0 + c1) + n * ho * wo * cout + wp.numel()
                  + (0 if residual is None else n * ho * wo * (1 if res_bcast else cout)))
    tag = f"conv{kh}x{kw}_s{sh}{sw}_d{dil}_cin{c0 + c1}_cout{cout}"
    rc = _run(tag, nbytes, lambda: _lib.load().codd_conv2d_nhwc(
        ctypes.byref(d), x.data_ptr(), None if x2 is None else x2.data_ptr(), wp.data_ptr(),
        None if bias is None else bias.data_ptr(), None if residual is None else residual.data_ptr(),
        out.data_ptr(), _stream()))
    _lib.check(rc, f"codd_conv2d_nhwc(k={kh}x{kw}, cin={c0}+{c1}, cout={cout})")
    return out


def conv3x3_image(left, right, wp, bias, cout):
    """First backbone conv on NCHW images; returns the NHWC feature map of cat([left, right])."""
    _require_cuda(left, right, wp, bias)
    left = left.contiguous()
    n, c, h, w = left.shape
    if c != 3:
        raise _lib.CoddError("codd_conv3x3_image expects 3-channel images")
    if right is not None:
        right = right.contiguous()
        if right.shape != left.shape:
            raise _lib.CoddError("left / right image shapes differ")
    nb = n if right is None else 2 * n
    out = empty_nhwc(nb, cout, h, w, left.device)
    rc = _run("conv3x3_image", 4 * nb * h * w * (3 + cout), lambda: _lib.load().codd_conv3x3_image(
        left.data_ptr(), None if right is None else right.data_ptr(), n, h, w, wp.data_ptr(), bias.data_ptr(), cout,
        out.data_ptr(), cout, _stream()))
    _lib.check(rc, "codd_conv3x3_image")
    return out


def upmerge_eligible(coarse, skip, cu, co):
    return (cu, co) in ((16, 16), (24, 24)) and coarse.shape[1] % 8 == 0 and skip.shape[1] % 8 == 0 and \
        skip.shape[2] == 2 * coarse.shape[2] and skip.shape[3] == 2 * coarse.shape[3]


def upmerge(coarse, skip, w_up, b_up, cu, w_merge, b_merge, co):
    """LeakyReLU(conv1x1(cat(skip, LeakyReLU(deconv2x2(coarse))))) in one kernel (backbone.py:17-32,75-88): the
    up-sampled tensor is never written.  w_up from pack_deconv_weight, w_merge from pack_conv_weight."""
    _require_cuda(coarse, skip, w_up, b_up, w_merge, b_merge)
    n, cs, h, w = skip.shape
    cc = coarse.shape[1]
    out = empty_nhwc(n, co, h, w, skip.device)
    nbytes = 4 * (n * h * w * (cs + co) + coarse.numel())
    rc = _run(f"upmerge_cc{cc}_cs{cs}_co{co}", nbytes, lambda: _lib.load().codd_upmerge_nhwc(
        coarse.data_ptr(), ld_of(coarse), cc, skip.data_ptr(), ld_of(skip), cs, w_up.data_ptr(), b_up.data_ptr(), cu,
        w_merge.data_ptr(), b_merge.data_ptr(), co, n, h, w, out.data_ptr(), ld_of(out), _stream()))
    _lib.check(rc, "codd_upmerge_nhwc")
    return out


def deconv2x2(x, wp, bias, cout, act=ACT_LEAKY):
    _require_cuda(x, wp, bias)
    n, cin, h, w = x.shape
    out = empty_nhwc(n, cout, 2 * h, 2 * w, x.device)
    rc = _run(f"deconv2x2_cin{cin}_cout{cout}", 4 * n * h * w * (cin + 4 * cout),
              lambda: _lib.load().codd_deconv2x2_nhwc(x.data_ptr(), ld_of(x), n, h, w, cin, wp.data_ptr(),
                                                      bias.data_ptr(), cout, out.data_ptr(), cout, act, _stream()))
    _lib.check(rc, "codd_deconv2x2_nhwc")
    return out


def tile_features(fea, w0p, b0, w1, b1, right):
    """K2.  fea [N,C,H,W] NHWC-backed -> PLANAR tile features [N,16,H/4,W/4] (left) or
    [N,16,H/4,W] (right: stride (4,1) over the input zero-padded 3 columns on the right)."""
    _require_cuda(fea, w0p, b0, w1, b1)
    n, c, h, w = fea.shape
    wo = w if right else w // 4
    out = torch.empty((n, 16, h // 4, wo), device=fea.device, dtype=torch.float32)
    flops_bytes = 4 * (n * h * w * c + out.numel())
    rc = _run("tile_features_" + ("right" if right else "left") + f"_c{c}", flops_bytes,
              lambda: _lib.load().codd_tile_features(fea.data_ptr(), ld_of(fea), c, n, h, w, w0p.data_ptr(),
                                                     b0.data_ptr(), w1.data_ptr(), b1.data_ptr(), 1 if right else 0,
                                                     out.data_ptr(), _stream()))
    _lib.check(rc, "codd_tile_features")
    return out


def tile_features_tc_eligible(fea):
    n, c, h, w = fea.shape
    return c in (16, 24, 32) and h % 4 == 0 and w % 4 == 0 and h * w >= 2048


def tile_features_tc(fea, w0split, b0, w1, b1, right):
    """K2 on the tensor cores (Cin = 16): same contract as tile_features, ``w0split`` from pack_conv_weight_tc4."""
    _require_cuda(fea, w0split, b0, w1, b1)
    n, c, h, w = fea.shape
    wo = w if right else w // 4
    out = torch.empty((n, 16, h // 4, wo), device=fea.device, dtype=torch.float32)
    rc = _run("tile_features_tc_" + ("right" if right else "left") + f"_c{c}", 4 * (n * h * w * c + out.numel()),
              lambda: _lib.load().codd_tile_features_tc(fea.data_ptr(), ld_of(fea), c, n, h, w, w0split.data_ptr(),
                                                        b0.data_ptr(), w1.data_ptr(), b1.data_ptr(), 1 if right else 0,
                                                        out.data_ptr(), _stream()))
    _lib.check(rc, "codd_tile_features_tc")
    return out


def cost_volume(tile_l, tile_r, max_disp, want_cv=False, want_argmin=True):
    """K1.  tile_l [N,16,h,w], tile_r [N,16,h,4w] (planar NCHW as K2 writes them; NHWC-backed
    inputs are transposed first).  Returns (cv or None, min_cost or None, min_disp or None);
    cv is [N,D,h,w] planar, min_* are [N,1,h,w]."""
    _require_cuda(tile_l, tile_r)
    tile_l, tile_r = planar(tile_l), planar(tile_r)
    n, c, h, w = tile_l.shape
    if c != 16 or tile_r.shape != (n, 16, h, 4 * w):
        raise _lib.CoddError(f"cost_volume expects [N,16,h,w] and [N,16,h,4w], got {tuple(tile_l.shape)} "
                             f"{tuple(tile_r.shape)}")
    dev = tile_l.device
    cv = torch.empty((n, max_disp, h, w), device=dev, dtype=torch.float32) if want_cv else None
    mc = torch.empty((n, 1, h, w), device=dev, dtype=torch.float32) if want_argmin else None
    md = torch.empty((n, 1, h, w), device=dev, dtype=torch.float32) if want_argmin else None
    nbytes = cost_volume_bytes(n, h, w, max_disp, want_cv, want_argmin)
    tag = "cost_volume_" + ("build" if want_cv else "") + ("argmin" if want_argmin else "")
    rc = _run(tag, nbytes, lambda: _lib.load().codd_cost_volume(
        tile_l.data_ptr(), tile_r.data_ptr(), n, h, w, max_disp,
        None if cv is None else cv.data_ptr(), None if mc is None else mc.data_ptr(),
        None if md is None else md.data_ptr(), _stream()))
    _lib.check(rc, "codd_cost_volume")
    return cv, mc, md


def cost_volume_pyramid(tiles, max_disps, want_cv=False, want_argmin=True):
    """K1 for all levels in one launch.  tiles: list of (tile_l [N,16,h,w], tile_r [N,16,h,4w]); max_disps: per level.
    Returns a list of (cv | None, min_cost | None, min_disp | None) per level, bit-identical to ops.cost_volume."""
    nl = len(tiles)
    tls = [planar(t[0]) for t in tiles]
    trs = [planar(t[1]) for t in tiles]
    _require_cuda(*tls, *trs)
    n = tls[0].shape[0]
    dev = tls[0].device
    outs, nbytes = [], 0
    for tl, tr, d in zip(tls, trs, max_disps):
        _, c, h, w = tl.shape
        if c != 16 or tr.shape != (n, 16, h, 4 * w):
            raise _lib.CoddError("cost_volume_pyramid expects [N,16,h,w] / [N,16,h,4w] pairs")
        cv = torch.empty((n, d, h, w), device=dev, dtype=torch.float32) if want_cv else None
        mc = torch.empty((n, 1, h, w), device=dev, dtype=torch.float32) if want_argmin else None
        md = torch.empty((n, 1, h, w), device=dev, dtype=torch.float32) if want_argmin else None
        outs.append((cv, mc, md))
        nbytes += cost_volume_bytes(n, h, w, d, want_cv, want_argmin)
    arr = lambda vals: (ctypes.c_void_p * nl)(*[None if v is None else v.data_ptr() for v in vals])
    iarr = lambda vals: (ctypes.c_int * nl)(*vals)
    tag = "cost_volume_pyramid_" + ("build" if want_cv else "") + ("argmin" if want_argmin else "")
    rc = _run(tag, nbytes, lambda: _lib.load().codd_cost_volume_pyramid(
        nl, arr(tls), arr(trs), n, iarr([t.shape[2] for t in tls]), iarr([t.shape[3] for t in tls]), iarr(list(max_disps)),
        arr([o[0] for o in outs]) if want_cv else None, arr([o[1] for o in outs]) if want_argmin else None,
        arr([o[2] for o in outs]) if want_argmin else None, _stream()))
    _lib.check(rc, "codd_cost_volume_pyramid")
    return outs


def cost_volume_bytes(n, h, w, max_disp, want_cv, want_argmin):
    """Algorithmic HBM bytes of K1: fp32 tile features in (16 ch x (w + 4w)
    columns), the volume out when materialised (4*D per tile), min cost + arg-min out (8 per tile)."""
    return n * h * w * (320 + (4 * max_disp if want_cv else 0) + (8 if want_argmin else 0))


def _is_planar(t):
    return t.is_contiguous() and not (t.shape[1] > 1 and t.stride(1) == 1)


def tile_hyp_init(min_cost, min_disp, feat, weight, bias):
    """feat may be NHWC-backed or planar NCHW (the K2 tile features)."""
    _require_cuda(min_cost, min_disp, feat, weight, bias)
    n, cf, h, w = feat.shape
    hyp = empty_nhwc(n, 16, h, w, feat.device)
    ldf = 0 if _is_planar(feat) else ld_of(feat)
    rc = _run("tile_hyp_init", 4 * n * h * w * (2 + cf + 16), lambda: _lib.load().codd_tile_hyp_init(
        min_cost.data_ptr(), min_disp.data_ptr(), feat.data_ptr(), ldf, cf, weight.data_ptr(), bias.data_ptr(),
        n, h, w, hyp.data_ptr(), 16, _stream()))
    _lib.check(rc, "codd_tile_hyp_init")
    return hyp


def plane_upsample(hyp, scale, size):
    _require_cuda(hyp)
    n, c, h, w = hyp.shape
    if c != 16:
        raise _lib.CoddError("plane_upsample expects 16-channel hypotheses")
    out = empty_nhwc(n, 16, h * size, w * size, hyp.device)
    rc = _run("plane_upsample", 64 * n * h * w * (1 + size * size), lambda: _lib.load().codd_plane_upsample(
        hyp.data_ptr(), ld_of(hyp), n, h, w, size, float(scale), out.data_ptr(), 16, _stream()))
    _lib.check(rc, "codd_plane_upsample")
    return out


def planar(t):
    """Plain contiguous NCHW copy of a feature map (what K4 wants for the right features);
    a tensor that already is contiguous NCHW is returned as is."""
    if _is_planar(t):
        return t
    return to_nchw(t)


# Right features of K4: read in place from the backbone's NHWC map (default; the CTA's window is staged into shared memory
# with 128-bit copies, no planar transpose) or from a planar copy (CODD_K4_NHWC=0, the earlier path).  tools/k4_probe.py at
# level 0, batch 8: smooth hypotheses 0.50 ms (NHWC) vs 0.51 ms (planar); hypotheses too scattered to stage 0.86 vs 0.69 ms
# (in-place 128-bit gathers are L1-latency bound); in the bench step both give 0.74 ms and NHWC saves the 5 transposes.
K4_NHWC = os.environ.get("CODD_K4_NHWC", "1") != "0"


def tile_warp_cost(fea_l, fea_r, cur, prev, dec_w, dec_b, want_raw=False, force_nhwc=False):
    """K4.  Returns aug [N,32|64,h,w] (and the raw 64-ch/set decrease input when want_raw).
    ``fea_r``: NHWC-backed (gathered in place) or plain contiguous NCHW (planar path)."""
    _require_cuda(fea_l, fea_r, cur, prev, dec_w, dec_b)
    nhwc_r = (K4_NHWC or force_nhwc) and not _is_planar(fea_r)
    if not nhwc_r:
        fea_r = planar(fea_r)
    n, c, H, W = fea_l.shape
    _, _, h, w = cur.shape
    if (H, W) != (4 * h, 4 * w) or fea_r.shape != fea_l.shape or cur.shape[1] != 16:
        raise _lib.CoddError("tile_warp_cost: feature / hypothesis shapes inconsistent")
    if prev is not None and prev.shape != (n, 16, h // 2, w // 2):
        raise _lib.CoddError("tile_warp_cost: previous-level hypotheses must be [N,16,h/2,w/2]")
    ca = 64 if prev is not None else 32
    aug = empty_nhwc(n, ca, h, w, cur.device)
    raw = empty_nhwc(n, 2 * ca, h, w, cur.device) if want_raw else None
    nbytes = tile_warp_bytes(n, c, h, w, prev is not None)
    tail = (cur.data_ptr(), ld_of(cur), None if prev is None else prev.data_ptr(), 0 if prev is None else ld_of(prev),
            dec_w.data_ptr(), dec_b.data_ptr(), n, h, w, aug.data_ptr(), ca, None if raw is None else raw.data_ptr(), _stream())
    tag = f"tile_warp_cost_c{c}_sets{2 if prev is not None else 1}"
    if nhwc_r:
        rc = _run(tag, nbytes, lambda: _lib.load().codd_tile_warp_cost_nhwc(
            fea_l.data_ptr(), ld_of(fea_l), fea_r.data_ptr(), ld_of(fea_r), c, *tail))
    else:
        rc = _run(tag, nbytes, lambda: _lib.load().codd_tile_warp_cost(
            fea_l.data_ptr(), ld_of(fea_l), fea_r.data_ptr(), c, *tail))
    _lib.check(rc, "codd_tile_warp_cost")
    return (aug, raw) if want_raw else aug


def tile_warp_bytes(n, c, h, w, has_prev):
    """Algorithmic HBM bytes of K4: both feature maps once (2*C per pixel, 16 pixels per tile),
    the hypotheses in (16 per tile, + the coarser level's 16 per 4 tiles) and the augmented
    hypothesis tensor out (32 or 64 per tile); fp32."""
    per_tile = 2 * c * 16 + 16 + (4 + 64 if has_prev else 32)
    return 4 * n * h * w * per_tile


def hyp_select(update, aug):
    _require_cuda(update, aug)
    n, cu, h, w = update.shape
    if cu != 34 or aug.shape != (n, 64, h, w):
        raise _lib.CoddError("hyp_select expects update [N,34,h,w] and aug [N,64,h,w]")
    out = empty_nhwc(n, 16, h, w, update.device)
    rc = _run("hyp_select", 4 * n * h * w * (18 + 16 + 16), lambda: _lib.load().codd_hyp_select(
        update.data_ptr(), ld_of(update), aug.data_ptr(), ld_of(aug), n, h, w, out.data_ptr(), 16, _stream()))
    _lib.check(rc, "codd_hyp_select")
    return out


# ----------------------------------------------------------------------------------------------
# Fusion (K13)
# ----------------------------------------------------------------------------------------------
def fusion_cues_lowres(feat_curr, feat_warp, fea_l, fea_r, pred_curr, pred_warp, ds=4, extra=None):
    """-> corr_feat [N,31,h,w] (NHWC, ld 32), disp2 [N,2,h,w] (NHWC)."""
    _require_cuda(feat_curr, feat_warp, fea_l, fea_r, pred_curr, pred_warp, extra)
    n, c, h, w = feat_curr.shape
    fea_r = planar(fea_r)
    pred_curr, pred_warp = pred_curr.contiguous(), pred_warp.contiguous()
    corr = empty_nhwc(n, 31, h, w, feat_curr.device, ld=32)
    disp2 = empty_nhwc(n, 2, h, w, feat_curr.device)
    cs = fea_l.shape[1]
    nbytes = 4 * n * h * w * (2 * c * 9 + 2 * cs + 2 + 34)
    rc = _run("fusion_cues_lowres", nbytes, lambda: _lib.load().codd_fusion_cues_lowres(
        feat_curr.data_ptr(), ld_of(feat_curr), feat_warp.data_ptr(), ld_of(feat_warp), fea_l.data_ptr(), ld_of(fea_l),
        fea_r.data_ptr(), cs, pred_curr.data_ptr(), pred_warp.data_ptr(), n, h, w, ds, corr.data_ptr(), 32,
        disp2.data_ptr(), 2, None if extra is None else extra.data_ptr(), 0 if extra is None else ld_of(extra),
        _stream()))
    _lib.check(rc, "codd_fusion_cues_lowres")
    return corr, disp2


def fusion_forget_in(pred_curr, pred_warp, flow_warp, conf_warp, weight, bias, want_cues=False):
    """-> forget_head[0] output [N,16,H,W] NHWC (and the 32 planar cues when want_cues)."""
    _require_cuda(pred_curr, pred_warp, flow_warp, conf_warp, weight, bias)
    n, _, h, w = pred_curr.shape
    pred_curr, pred_warp = pred_curr.contiguous(), pred_warp.contiguous()
    flow_warp, conf_warp = planar(flow_warp), planar(conf_warp)
    out = empty_nhwc(n, 16, h, w, pred_curr.device)
    cues = torch.empty((n, 32, h, w), device=pred_curr.device) if want_cues else None
    rc = _run("fusion_forget_in", 4 * n * h * w * (8 + 16), lambda: _lib.load().codd_fusion_forget_in(
        pred_curr.data_ptr(), pred_warp.data_ptr(), flow_warp.data_ptr(), conf_warp.data_ptr(), weight.data_ptr(),
        bias.data_ptr(), n, h, w, out.data_ptr(), 16, None if cues is None else cues.data_ptr(), _stream()))
    _lib.check(rc, "codd_fusion_forget_in")
    return (out, cues) if want_cues else out


def fusion_blend(pred_curr, pred_warp, r8, weight, bias, wf_lowres, ds=4):
    """-> (disp_fused, fusion_weights, reset_weights), each [N,1,H,W]."""
    _require_cuda(pred_curr, pred_warp, r8, weight, bias, wf_lowres)
    n, _, h, w = pred_curr.shape
    pred_curr, pred_warp, wf_lowres = pred_curr.contiguous(), pred_warp.contiguous(), wf_lowres.contiguous()
    dev = pred_curr.device
    fused, wf, wr = (torch.empty((n, 1, h, w), device=dev) for _ in range(3))
    rc = _run("fusion_blend", 4 * n * h * w * (2 + 8 + 3), lambda: _lib.load().codd_fusion_blend(
        pred_curr.data_ptr(), pred_warp.data_ptr(), r8.data_ptr(), ld_of(r8), weight.data_ptr(), bias.data_ptr(),
        wf_lowres.data_ptr(), n, h, w, ds, fused.data_ptr(), wf.data_ptr(), wr.data_ptr(), _stream()))
    _lib.check(rc, "codd_fusion_blend")
    return fused, wf, wr


# ----------------------------------------------------------------------------------------------
# Motion / RAFT3D non-convolutional ops (K9-K12)
# ----------------------------------------------------------------------------------------------
def raft_motion_info(Ts, depth1, depth2_inv, intr):
    """-> (coords1_xyz [N,h,w,3], motion_info [N,9,h,w] NHWC-backed, ld 12)."""
    _require_cuda(Ts, depth1, depth2_inv, intr)
    n, h, w = depth1.shape
    Ts, depth1, depth2_inv, intr = Ts.contiguous(), depth1.contiguous(), depth2_inv.contiguous(), intr.contiguous()
    xyz = torch.empty((n, h, w, 3), device=Ts.device)
    info = empty_nhwc(n, 9, h, w, Ts.device, ld=12)
    rc = _run("raft_motion_info", 4 * n * h * w * (7 + 2 + 3 + 9), lambda: _lib.load().codd_raft_motion_info(
        Ts.data_ptr(), depth1.data_ptr(), depth2_inv.data_ptr(), intr.data_ptr(), n, h, w, xyz.data_ptr(),
        info.data_ptr(), 12, _stream()))
    _lib.check(rc, "codd_raft_motion_info")
    return xyz, info


def avgpool2(x):
    _require_cuda(x)
    n, c, h, w = x.shape
    out = empty_nhwc(n, c, h // 2, w // 2, x.device)
    rc = _run("avgpool2", 5 * n * c * h * w, lambda: _lib.load().codd_avgpool2_nhwc(
        x.data_ptr(), ld_of(x), n, h, w, c, out.data_ptr(), c, _stream()))
    _lib.check(rc, "codd_avgpool2_nhwc")
    return out


def corr_pyramid(fmap2, levels=4):
    """fmap2 NHWC-backed -> list of pooled maps (level 0 = fmap2 itself)."""
    pyr = [to_nhwc(fmap2)]
    for _ in range(levels - 1):
        pyr.append(avgpool2(pyr[-1]))
    return pyr


def corr_lookup(fmap1, pyramid, coords_xyz, radius=3):
    """fmap1 [N,C,h,w] NHWC-backed, pyramid from corr_pyramid, coords_xyz [N,h,w,>=2] -> [N,L*(2r+1)^2,h,w] NHWC."""
    _require_cuda(fmap1, coords_xyz)
    n, c, h, w = fmap1.shape
    levels = len(pyramid)
    nch = levels * (2 * radius + 1) ** 2
    out = empty_nhwc(n, nch, h, w, fmap1.device)
    ptrs = (ctypes.c_void_p * levels)(*[p.data_ptr() for p in pyramid])
    lds = (ctypes.c_int * levels)(*[ld_of(p) for p in pyramid])
    coords_xyz = coords_xyz.contiguous()
    rc = _run("corr_lookup", 4 * n * h * w * (c + nch), lambda: _lib.load().codd_corr_lookup(
        fmap1.data_ptr(), ld_of(fmap1), ptrs, lds, levels, coords_xyz.data_ptr(), coords_xyz.shape[-1], n, h, w, c,
        radius, out.data_ptr(), nch, _stream()))
    _lib.check(rc, "codd_corr_lookup")
    return out


def se3_gn_step(Ts, ae, target, weight, depth, intr, radius=32, lm=1e-4, ep=10.0):
    """Ts [N,h,w,7]; ae [N,32,h,w], target / weight [N,3,h,w] NHWC-backed -> new Ts."""
    _require_cuda(Ts, ae, target, weight, depth, intr)
    n, h, w = depth.shape
    Ts, depth, intr = Ts.contiguous(), depth.contiguous(), intr.contiguous()
    out = torch.empty_like(Ts)
    rc = _run("se3_gn_step", 4 * n * h * w * (7 + 32 + 3 + 3 + 1 + 7), lambda: _lib.load().codd_se3_gn_step(
        Ts.data_ptr(), ae.data_ptr(), ld_of(ae), target.data_ptr(), ld_of(target), weight.data_ptr(), ld_of(weight),
        depth.data_ptr(), intr.data_ptr(), n, h, w, radius, lm, ep, out.data_ptr(), _stream()))
    _lib.check(rc, "codd_se3_gn_step")
    return out


def cvx_upsample(data, mask):
    """data [N,h,w,dim] contiguous, mask [N,576,h,w] NHWC-backed -> [N,8h,8w,dim]."""
    _require_cuda(data, mask)
    n, h, w, dim = data.shape
    data = data.contiguous()
    out = torch.empty((n, 8 * h, 8 * w, dim), device=data.device)
    rc = _run("cvx_upsample", 4 * n * h * w * (dim + 576 + 64 * dim), lambda: _lib.load().codd_cvx_upsample(
        data.data_ptr(), dim, dim, mask.data_ptr(), ld_of(mask), n, h, w, out.data_ptr(), dim, _stream()))
    _lib.check(rc, "codd_cvx_upsample")
    return out


def se3_upsample_flow(Ts, mask, depth, intr):
    """-> (Ts_up [N,8h,8w,7], flow [N,8h,8w,3])."""
    _require_cuda(Ts, mask, depth, intr)
    n, h, w, _ = Ts.shape
    Ts, depth, intr = Ts.contiguous(), depth.contiguous(), intr.contiguous()
    ws = torch.empty((n, h, w, 6), device=Ts.device)
    up = torch.empty((n, 8 * h, 8 * w, 7), device=Ts.device)
    flow = torch.empty((n, 8 * h, 8 * w, 3), device=Ts.device)
    rc = _run("se3_upsample_flow", 4 * n * h * w * (7 + 576 + 64 * 11), lambda: _lib.load().codd_se3_upsample_flow(
        Ts.data_ptr(), mask.data_ptr(), ld_of(mask), depth.data_ptr(), intr.data_ptr(), n, h, w, ws.data_ptr(),
        up.data_ptr(), flow.data_ptr(), _stream()))
    _lib.check(rc, "codd_se3_upsample_flow")
    return up, flow


def splat_warp(Ts, depth, intr, feat, radius, bf=0.0, want_disp=False):
    """feat [N,C,h,w] NHWC-backed -> (warped [N,C,h,w] NHWC, zbuf [N,1,h,w], disp [N,1,h,w] | None)."""
    _require_cuda(Ts, depth, intr, feat)
    n, c, h, w = feat.shape
    Ts, depth, intr = Ts.contiguous(), depth.contiguous(), intr.contiguous()
    out = empty_nhwc(n, c, h, w, feat.device)
    zbuf = torch.empty((n, 1, h, w), device=feat.device)
    disp = torch.empty((n, 1, h, w), device=feat.device) if want_disp else None
    nbytes = _lib.load().codd_splat_workspace_bytes(n, h, w)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=feat.device)
    LAUNCHES[0] += 2
    rc = _run("splat_warp", 4 * n * h * w * (7 + 1 + 2 * c + 2), lambda: _lib.load().codd_splat_warp(
        Ts.data_ptr(), depth.data_ptr(), intr.data_ptr(), feat.data_ptr(), ld_of(feat), c, n, h, w, float(radius),
        float(bf), out.data_ptr(), c, zbuf.data_ptr(), None if disp is None else disp.data_ptr(), ws.data_ptr(), nbytes,
        _stream()))
    _lib.check(rc, "codd_splat_warp")
    return out, zbuf, disp


# ----------------------------------------------------------------------------------------------
# RAFT3D network glue (instance norm, bilinear resize, element-wise, depth conversion, sub-sampling)
# ----------------------------------------------------------------------------------------------
def instance_norm(x, relu=True, residual=None, eps=1e-5):
    """InstanceNorm2d(affine=False) (+ReLU); with ``residual``: relu(residual + y) (ResidualBlock tail)."""
    _require_cuda(x, residual)
    n, c, h, w = x.shape
    out = empty_nhwc(n, c, h, w, x.device)
    nb = _lib.load().codd_instance_norm_workspace_bytes(n, c)
    ws = torch.empty(nb, dtype=torch.uint8, device=x.device)
    LAUNCHES[0] += 1
    rc = _run("instance_norm", 4 * x.numel() * (3 + (0 if residual is None else 1)),
              lambda: _lib.load().codd_instance_norm_nhwc(
                  x.data_ptr(), ld_of(x), n, h, w, c, eps, 1 if relu else 0,
                  None if residual is None else residual.data_ptr(), 0 if residual is None else ld_of(residual),
                  out.data_ptr(), c, ws.data_ptr(), nb, _stream()))
    _lib.check(rc, "codd_instance_norm_nhwc")
    return out


def resize_bilinear(x, size, align_corners, base=None, relu=False, out=None):
    """out = relu?(base + F.interpolate(x, size, mode='bilinear', align_corners))."""
    _require_cuda(x, base, out)
    n, c, h, w = x.shape
    ho, wo = size
    if out is None:
        out = empty_nhwc(n, c, ho, wo, x.device)
    rc = _run("resize_bilinear", 4 * (x.numel() + n * c * ho * wo * (1 if base is None else 2)),
              lambda: _lib.load().codd_resize_bilinear_nhwc(
                  x.data_ptr(), ld_of(x), n, h, w, c, None if base is None else base.data_ptr(),
                  0 if base is None else ld_of(base), out.data_ptr(), ld_of(out), ho, wo, 1 if align_corners else 0,
                  1 if relu else 0, _stream()))
    _lib.check(rc, "codd_resize_bilinear_nhwc")
    return out


EW_ACT, EW_MUL, EW_GRU, EW_ADD_ACT, EW_RECIP = 0, 1, 2, 3, 4


def eltwise(op, a, b=None, c=None, act=ACT_NONE, out=None):
    """Element-wise glue on NHWC-backed logical-NCHW tensors (see codd_eltwise_nhwc)."""
    _require_cuda(a, b, c, out)
    n, ch, h, w = a.shape
    if out is None:
        out = empty_nhwc(n, ch, h, w, a.device)
    nin = 1 + (b is not None) + (c is not None)
    rc = _run(f"eltwise_op{op}", 4 * a.numel() * (nin + 1), lambda: _lib.load().codd_eltwise_nhwc(
        op, act, a.data_ptr(), ld_of(a), None if b is None else b.data_ptr(), 0 if b is None else ld_of(b),
        None if c is None else c.data_ptr(), 0 if c is None else ld_of(c), out.data_ptr(), ld_of(out), n * h * w, ch,
        _stream()))
    _lib.check(rc, "codd_eltwise_nhwc")
    return out


def disp_to_depth(disp, bf):
    """clip(bf / (disp + 1e-5), 0, bf) on a contiguous tensor of any shape."""
    _require_cuda(disp)
    disp = disp.contiguous()
    out = torch.empty_like(disp)
    rc = _run("disp_to_depth", 8 * disp.numel(), lambda: _lib.load().codd_disp_to_depth(
        disp.data_ptr(), disp.numel(), float(bf), out.data_ptr(), _stream()))
    _lib.check(rc, "codd_disp_to_depth")
    return out


def subsample(x, offset, stride, recip=False):
    """x [N,H,W,C] contiguous (channels last) -> x[:, offset::stride, offset::stride, :] (optionally 1/x)."""
    _require_cuda(x)
    x = x.contiguous()
    n, h, w, c = x.shape
    ho, wo = (h - offset + stride - 1) // stride, (w - offset + stride - 1) // stride
    out = torch.empty((n, ho, wo, c), device=x.device)
    rc = _run("subsample", 8 * out.numel(), lambda: _lib.load().codd_subsample_nhwc(
        x.data_ptr(), c, n, h, w, c, offset, stride, 1 if recip else 0, out.data_ptr(), c, _stream()))
    _lib.check(rc, "codd_subsample_nhwc")
    return out


# ----------------------------------------------------------------------------------------------
# N1 input staging: Normalize + Pad(size_divisor) + HWC->CHW of the reference's test pipeline on the GPU
# ----------------------------------------------------------------------------------------------
IMG_NORM = dict(mean=(123.675, 116.28, 103.53), std=(58.395, 57.12, 57.375), to_rgb=True)   # configs/datasets/*.py


def stage_images_u8(img, mean=IMG_NORM["mean"], std=IMG_NORM["std"], to_rgb=True, size_divisor=64):
    """img: uint8 CUDA tensor [N,H,W,3] (as cv2 / mmcv load frames) -> normalised fp32 [N,3,Hp,Wp], Hp/Wp = H/W rounded up
    to ``size_divisor``, reflect-padded on the bottom / right (datasets/transforms.py:147-176, 391-421)."""
    if not img.is_cuda or img.dtype != torch.uint8 or img.dim() != 4 or img.shape[-1] != 3:
        raise _lib.CoddError("stage_images_u8 expects a uint8 CUDA tensor [N,H,W,3]")
    img = img.contiguous()
    n, h, w, _ = img.shape
    hp = -(-h // size_divisor) * size_divisor
    wp = -(-w // size_divisor) * size_divisor
    out = torch.empty((n, 3, hp, wp), device=img.device, dtype=torch.float32)
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    rc = _run("stage_images_u8", img.numel() + 4 * out.numel(), lambda: _lib.load().codd_stage_images_u8(
        img.data_ptr(), n, h, w, m, s, 1 if to_rgb else 0, hp, wp, out.data_ptr(), _stream()))
    _lib.check(rc, "codd_stage_images_u8")
    return out


# ------------------------------------------------------------------------------------------
# N3: on-GPU evaluation (codd.py:435-517).  Accumulator rows live on the device; nothing here synchronises.
# ------------------------------------------------------------------------------------------
def _plane_view(t, name):
    """[N,1,H,W] fp32 CUDA tensor whose rows are contiguous (e.g. the [:h,:w] crop of a padded map)
    -> (data_ptr, sample stride, row stride)."""
    if t.dim() != 4 or t.shape[1] != 1 or (t.shape[3] > 1 and t.stride(3) != 1):
        raise _lib.CoddError(f"{name}: expected an [N,1,H,W] tensor with contiguous rows, got {tuple(t.shape)} / {t.stride()}")
    return t.data_ptr(), (t.stride(0) if t.shape[0] > 1 else t.shape[2] * t.stride(2)), t.stride(2)


def disp_metrics(pred, gt, disp_range, acc, seg=None, mask_out=None):
    """acc[0..3] += (#valid, sum |pred-gt|, #(|pred-gt| > 3), #(gt > 0)); mask_out (uint8 [N,1,H,W]) = validity."""
    _require_cuda(pred, gt, seg)
    n, _, h, w = gt.shape
    gt = gt.contiguous()
    seg = None if seg is None else seg.contiguous()
    pp, pss, prs = _plane_view(pred[:, :, :h, :w], "pred")
    rc = _run("disp_metrics", 4 * n * h * w * (2 + (seg is not None)) + n * h * w, lambda: _lib.load().codd_disp_metrics(
        pp, pss, prs, gt.data_ptr(), None if seg is None else seg.data_ptr(), n, h, w, float(disp_range[0]),
        float(disp_range[1]), None if mask_out is None else mask_out.data_ptr(), acc.data_ptr(), _stream()))
    _lib.check(rc, "codd_disp_metrics")
    return acc


def temporal_metrics(flow_prev, gt, pred, gt_prev, pred_prev, mask_prev, disp_range, acc, seg=None, gt_disp2_prev=None,
                     gt_pos_count=None):
    """acc[0..8] += the temporal-EPE sums of one frame pair (see include/codd_b200.h)."""
    _require_cuda(flow_prev, gt, pred, gt_prev, pred_prev, seg, gt_disp2_prev)
    n, _, h, w = gt.shape
    flow_prev, gt, gt_prev = flow_prev.contiguous(), gt.contiguous(), gt_prev.contiguous()
    seg = None if seg is None else seg.contiguous()
    g2 = None if gt_disp2_prev is None else gt_disp2_prev.contiguous()
    pp, pss, prs = _plane_view(pred[:, :, :h, :w], "pred")
    qp, qss, qrs = _plane_view(pred_prev[:, :, :h, :w], "pred_prev")
    rc = _run("temporal_metrics", 4 * n * h * w * 8, lambda: _lib.load().codd_temporal_metrics(
        flow_prev.data_ptr(), gt.data_ptr(), pp, pss, prs, None if seg is None else seg.data_ptr(), gt_prev.data_ptr(),
        qp, qss, qrs, mask_prev.data_ptr(), None if g2 is None else g2.data_ptr(),
        None if gt_pos_count is None else gt_pos_count.data_ptr(), n, h, w, float(disp_range[0]), float(disp_range[1]),
        acc.data_ptr(), _stream()))
    _lib.check(rc, "codd_temporal_metrics")
    return acc


def sceneflow_metrics(Ts, pred_prev, intrinsics, flow_prev, gt_disp_change, gt_prev, disp_range, acc, seg=None,
                      flow_occ=None):
    """acc[0..4] += (#valid, sum scene-flow EPE, sum optical-flow EPE, #(sf < 1), #(of < 1)) of one frame pair
    (codd.py:519-575).  Ts: [N,H',W',7] dense SE3 field (cropped to gt's size here); intrinsics [N,4] CUDA."""
    _require_cuda(Ts, pred_prev, intrinsics, flow_prev, gt_disp_change, gt_prev, seg)
    n, _, h, w = gt_prev.shape
    if Ts.dim() != 4 or Ts.shape[-1] != 7 or Ts.stride(3) != 1 or Ts.stride(2) != 7:
        raise _lib.CoddError("sceneflow_metrics: Ts must be [N,H,W,7] with contiguous pixels")
    flow_prev, gt_disp_change, gt_prev = flow_prev.contiguous(), gt_disp_change.contiguous(), gt_prev.contiguous()
    intrinsics = intrinsics.contiguous()
    seg = None if seg is None else seg.contiguous()
    occ = None if flow_occ is None else flow_occ.to(torch.uint8).contiguous()
    qp, qss, qrs = _plane_view(pred_prev[:, :, :h, :w], "pred_prev")
    tss = Ts.stride(0) if Ts.shape[0] > 1 else Ts.shape[1] * Ts.stride(1)
    rc = _run("sceneflow_metrics", 4 * n * h * w * 14, lambda: _lib.load().codd_sceneflow_metrics(
        Ts.data_ptr(), tss, Ts.stride(1), qp, qss, qrs, intrinsics.data_ptr(), flow_prev.data_ptr(),
        gt_disp_change.data_ptr(), gt_prev.data_ptr(), None if seg is None else seg.data_ptr(),
        None if occ is None else occ.data_ptr(), n, h, w, float(disp_range[0]), float(disp_range[1]), acc.data_ptr(),
        _stream()))
    _lib.check(rc, "codd_sceneflow_metrics")
    return acc


def gt_disp_change(flow_prev, gt_curr, gt_prev, flow_occ_prev=None):
    """utils/misc.py:39-59: (change [N,1,H,W], warped gt [N,1,H,W]) for one frame pair."""
    _require_cuda(flow_prev, gt_curr, gt_prev)
    n, _, h, w = gt_prev.shape
    flow_prev, gt_curr, gt_prev = flow_prev.contiguous(), gt_curr.contiguous(), gt_prev.contiguous()
    occ = None if flow_occ_prev is None else flow_occ_prev.to(torch.uint8).contiguous()
    change = torch.empty_like(gt_prev)
    warped = torch.empty_like(gt_prev)
    rc = _run("gt_disp_change", 4 * n * h * w * 6, lambda: _lib.load().codd_gt_disp_change(
        flow_prev.data_ptr(), gt_curr.data_ptr(), gt_prev.data_ptr(), None if occ is None else occ.data_ptr(), n, h, w,
        change.data_ptr(), warped.data_ptr(), _stream()))
    _lib.check(rc, "codd_gt_disp_change")
    return change, warped
