"""CPU oracle for the HITNetMF stereo hot path.

TEST INFRASTRUCTURE ONLY.  This file is the *checker*: a CPU (torch fp32 / numpy)
restatement of the reference algorithm, written functionally over a flat
``state_dict`` whose keys are the reference's own parameter names.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product path (``codd_b200``) never does.

Parity pin: every function below is compared against the unmodified reference file it
restates (``tests/test_oracle_vs_reference.py``, run where ``/root/reference`` is
mounted) and against the committed fixtures in ``tests/golden/`` that were generated
by executing the reference (``oracle/gen_golden.py``).  The reference ships no golden
vectors of its own, so those two are the pin.

Arithmetic notes (all fp32, no FMA contraction on the discrete-decision paths):
  * cost volume: integer gather at ``4*j - d`` with zero fill, channel sum sequential
    c = 0..C-1  -> bitwise equal to the reference's grid_sample(nearest)+norm(p=1) on CPU.
  * plane up-sampling: ``(d + a*dx) + b*dy`` then ``* scale``.
  * right-feature warp: the reference normalises the sample grid to [-1, 1] and torch
    un-normalises it again; both roundings are restated (``warp_coords``), because the
    resulting sub-ulp offsets decide which two texels / rows are blended.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LEAKY = 0.2


# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------
def _lrelu(x):
    return F.leaky_relu(x, LEAKY)


def _conv(sd, name, x, stride=1, padding=0, dilation=1):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride,
                    padding=padding, dilation=dilation)


def _sub(sd, prefix):
    """state_dict view with ``prefix`` stripped."""
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


# --------------------------------------------------------------------------------------
# a1  HITUNet.forward                      (model/stereo/hitnet/backbone.py:69-88)
# --------------------------------------------------------------------------------------
def backbone(sd, img):
    """5-level U-Net.  Returns [1/16 x32ch, 1/8 x24, 1/4 x24, 1/2 x16, 1/1 x16]."""

    def down(name, x):  # conv_down, backbone.py:8-14
        x = _lrelu(_conv(sd, name + ".0", x, stride=2, padding=1))
        return _lrelu(_conv(sd, name + ".2", x, padding=1))

    def up(name, x):  # conv_up, backbone.py:17-21
        return _lrelu(F.conv_transpose2d(x, sd[name + ".0.weight"], sd[name + ".0.bias"], stride=2))

    def merge(name, skip, x):  # conv_merge, backbone.py:24-32 (cat order: skip first, :76)
        x = torch.cat((skip, x), 1)
        x = _lrelu(_conv(sd, name + ".0", x))
        x = _lrelu(_conv(sd, name + ".2", x, padding=1))
        return _lrelu(_conv(sd, name + ".4", x, padding=1))

    x0 = _lrelu(_conv(sd, "conv1.0", img, padding=1))
    x1 = down("down1", x0)
    x2 = down("down2", x1)
    x3 = down("down3", x2)
    x4 = down("down4.0", x3)
    x4 = _lrelu(_conv(sd, "down4.1", x4, padding=1))
    x4 = _lrelu(_conv(sd, "down4.3", x4, padding=1))
    u4 = merge("merge4", x3, up("up4", x4))
    u3 = merge("merge3", x2, up("up3", u4))
    u2 = merge("merge2", x1, up("up2", u3))
    u1 = merge("merge1", x0, up("up1", u2))
    return [x4, u4, u3, u2, u1]


# --------------------------------------------------------------------------------------
# a2  TileInitialization.tile_features      (model/stereo/hitnet/initialization.py:119-156)
# --------------------------------------------------------------------------------------
_TILE_CONV = ["tile_conv16x", "tile_conv8x", "tile_conv4x", "tile_conv2x", "tile_conv1x"]
_TILE_DSC = ["tile_fea_dscrpt16x", "tile_fea_dscrpt8x", "tile_fea_dscrpt4x",
             "tile_fea_dscrpt2x", "tile_fea_dscrpt1x"]


def tile_features_level(sd, name, fea_l, fea_r):
    """Left: 4x4 stride-4 conv; right: same weights, stride (4,1) on the input padded by 3
    zero columns on the right, so right column x sees input columns x..x+3."""
    tl = _lrelu(_conv(sd, name + ".0", fea_l, stride=4))
    tl = _lrelu(_conv(sd, name + ".2", tl))
    tr = _lrelu(_conv(sd, name + ".0", F.pad(fea_r, (0, 3, 0, 0)), stride=(4, 1)))
    tr = _lrelu(_conv(sd, name + ".2", tr))
    return tl, tr


def tile_features(sd, fea_l, fea_r):
    """fea_* are coarse->fine pyramids (backbone output order)."""
    return [tile_features_level(sd, _TILE_CONV[k], fea_l[k], fea_r[k]) for k in range(5)]


# --------------------------------------------------------------------------------------
# a3  calc_init_disp                         (model/stereo/hitnet/initialization.py:18-45)
# --------------------------------------------------------------------------------------
def cost_volume(tile_l, tile_r, max_disp):
    """cv[n,d,i,j] = sum_c |L[n,c,i,j] - R[n,c,i,4j-d]|, R taken as 0 when 4j-d < 0.

    The reference builds this with a 5-D nearest grid_sample (integer positions, so the
    sample is an exact gather) followed by ``torch.norm(p=1, dim=1)``, whose CPU kernel
    adds channels sequentially in fp32.  Restated with an explicit gather; the channel loop
    is sequential so the result is bit-identical.
    """
    n, c, h, w = tile_l.shape
    wr = tile_r.shape[3]
    j4 = torch.arange(w, device=tile_l.device) * 4
    out = torch.empty(n, max_disp, h, w, dtype=torch.float32, device=tile_l.device)
    for d in range(max_disp):
        x = j4 - d
        ok = (x >= 0) & (x <= wr - 1)
        g = tile_r[:, :, :, x.clamp(0, wr - 1)] * ok.to(tile_r.dtype)
        diff = (tile_l - g).abs()
        acc = diff[:, 0].clone()
        for ch in range(1, c):
            acc = acc + diff[:, ch]
        out[:, d] = acc
    return out


def cost_volume_reference_form(tile_l, tile_r, max_disp):
    """The same volume computed the way the reference computes it (initialization.py:18-45):
    a [N,D,h,w,3] sampling grid, one 5-D nearest ``grid_sample`` that materialises the
    [N,C,D,h,w] shifted right features, subtract, L1 norm.  Bit-identical to ``cost_volume``;
    kept because it is what the reference's CPU time is made of (~50x the memory traffic), so the
    CPU baseline / ``--impl reference`` arm of bench.py times this form, not the cheap gather."""
    n, c, h, w = tile_l.shape
    wr, hr = tile_r.shape[3], tile_r.shape[2]
    dev = tile_l.device
    xs = torch.arange(0, wr, 4, dtype=torch.float32, device=dev)[:w] / (wr - 1) * 2 - 1          # columns 4j
    ys = torch.arange(h, dtype=torch.float32, device=dev) / (hr - 1) * 2 - 1
    shift = torch.arange(max_disp, dtype=torch.float32, device=dev) / (wr - 1) * 2
    gx = xs.view(1, 1, 1, w) - shift.view(1, max_disp, 1, 1)
    grid = torch.stack((gx.expand(n, max_disp, h, w), ys.view(1, 1, h, 1).expand(n, max_disp, h, w),
                        torch.zeros(n, max_disp, h, w, device=dev)), -1)
    shifted = F.grid_sample(tile_r.unsqueeze(2), grid, mode="nearest", align_corners=True, padding_mode="zeros")
    return torch.norm(tile_l.unsqueeze(2) - shifted, p=1, dim=1)


def cost_volume_numpy(tile_l, tile_r, max_disp):
    """Same as ``cost_volume`` on numpy arrays (independent second statement)."""
    tl = np.asarray(tile_l, dtype=np.float32)
    tr = np.asarray(tile_r, dtype=np.float32)
    n, c, h, w = tl.shape
    wr = tr.shape[3]
    pad = np.concatenate([np.zeros((n, c, h, max_disp), np.float32), tr], axis=3)
    out = np.empty((n, max_disp, h, w), np.float32)
    cols = np.arange(w) * 4 + max_disp
    for d in range(max_disp):
        g = pad[:, :, :, cols - d]
        acc = np.abs(tl[:, 0] - g[:, 0])
        for ch in range(1, c):
            acc = (acc + np.abs(tl[:, ch] - g[:, ch])).astype(np.float32)
        out[:, d] = acc
    assert wr >= 4 * (w - 1) + 1
    return out


def cost_volume_argmin(tile_l, tile_r, max_disp):
    """(min cost, first arg-min) of ``cost_volume`` — initialization.py:167-171."""
    cv = cost_volume(tile_l, tile_r, max_disp)
    cost, idx = torch.min(cv, 1)
    return cost, idx


# --------------------------------------------------------------------------------------
# a4  tile_hypothesis_pyramid                (model/stereo/hitnet/initialization.py:158-225)
# --------------------------------------------------------------------------------------
def tile_hypotheses(sd, tile_pyr, fea_l, max_disp, return_cv=False, reference_form=False):
    """16-channel hypotheses [d, dx=0, dy=0, 13-ch descriptor] per level, coarse->fine.

    Descriptor input is cat[min cost, feature]; the feature is the left *tile* feature for
    the two coarsest levels and the backbone pyramid entries 0,1,2 for the finer three
    (initialization.py:186-190; resolutions coincide with the tile grids)."""
    hyps, cvs = [], []
    for k in range(5):
        tl, tr = tile_pyr[k]
        cv = (cost_volume_reference_form if reference_form else cost_volume)(tl, tr, max_disp // (16 >> k))
        cost, idx = torch.min(cv, 1)
        feat = tl if k < 2 else fea_l[k - 2]
        dsc = _lrelu(_conv(sd, _TILE_DSC[k] + ".0", torch.cat([cost.unsqueeze(1), feat], 1)))
        d = idx.float().unsqueeze(1)
        z = torch.zeros_like(d)
        hyps.append(torch.cat([d, z, z, dsc], 1))
        cvs.append(cv)
    return (hyps, cvs) if return_cv else hyps


# --------------------------------------------------------------------------------------
# a5  to_plane / upsample                    (model/stereo/hitnet/propagation.py:10-32)
# --------------------------------------------------------------------------------------
def plane_offsets(size, device=None):
    """linspace(-(s-1)/2, (s-1)/2, s): exactly representable half-integers."""
    return torch.arange(size, dtype=torch.float32, device=device) - (size - 1) / 2.0


def plane_expand(d, dx, dy, size):
    """[N,1,h,w] -> [N,1,h*size,w*size]:  (d + c[x%size]*dx) + c[y%size]*dy."""
    c = plane_offsets(size, d.device)
    n, _, h, w = d.shape
    a = c.view(1, 1, 1, size).repeat(1, 1, h * size, w)           # varies with x
    b = c.view(1, 1, size, 1).repeat(1, 1, h, w * size)           # varies with y
    rep = lambda t: t.repeat_interleave(size, 2).repeat_interleave(size, 3)
    return rep(d) + a * rep(dx) + b * rep(dy)


def plane_upsample(hyp, scale=2, size=2):
    d = plane_expand(hyp[:, 0:1], hyp[:, 1:2], hyp[:, 2:3], size) * scale
    rest = hyp[:, 1:].repeat_interleave(size, 2).repeat_interleave(size, 3)
    return torch.cat((d, rest), 1)


# --------------------------------------------------------------------------------------
# a6  warp                                   (model/stereo/hitnet/propagation.py:35-58)
# --------------------------------------------------------------------------------------
def warp_coords(disp):
    """Pixel-space sample coordinates (ix [N,H,W], iy [H]) exactly as the reference +
    torch compute them: normalise ``2*(x-d)/(W-1) - 1`` (true division), then
    grid_sample's align_corners=True un-normalisation ``((g+1)/2)*(W-1)``."""
    n, _, h, w = disp.shape
    xs = torch.arange(w, dtype=torch.float32, device=disp.device).view(1, 1, w)
    ys = torch.arange(h, dtype=torch.float32, device=disp.device)
    gx = 2.0 * (xs - disp[:, 0]) / max(w - 1, 1) - 1.0
    gy = 2.0 * ys / max(h - 1, 1) - 1.0
    ix = ((gx + 1.0) / 2.0) * (w - 1)
    iy = ((gy + 1.0) / 2.0) * (h - 1)
    return ix, iy


def warp_right(fea_r, disp):
    """Bilinear, zeros padding, sampled at (x - disp, y).  Reference form (grid_sample)."""
    n, c, h, w = fea_r.shape
    xs = torch.arange(w, dtype=torch.float32, device=disp.device).view(1, 1, w).expand(n, h, w)
    ys = torch.arange(h, dtype=torch.float32, device=disp.device).view(1, h, 1).expand(n, h, w)
    gx = 2.0 * (xs - disp[:, 0]) / max(w - 1, 1) - 1.0
    gy = 2.0 * ys / max(h - 1, 1) - 1.0
    return F.grid_sample(fea_r, torch.stack((gx, gy), -1), mode="bilinear",
                         padding_mode="zeros", align_corners=True)


def warp_right_direct(fea_r, disp):
    """Explicit-arithmetic statement of ``warp_right`` (what the CUDA kernel mirrors).

    torch's CPU bilinear kernel: w = ix - floor(ix), e = 1 - w, n = iy - floor(iy), s = 1 - n;
    out = fma(se, n*w, fma(sw, n*e, fma(ne, s*w, nw*(s*e)))); out-of-range taps contribute 0.
    (Found by matching F.grid_sample bit-for-bit; the fused multiply-adds are emulated here
    in float64, where a 24x24-bit product is exact.)
    """
    n, c, h, w = fea_r.shape
    ix, iy = warp_coords(disp)
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    fw = ix - x0
    fe = 1.0 - fw
    fn = (iy - y0).view(1, h, 1)
    fs = 1.0 - fn
    x0 = x0.long()
    y0 = y0.long().view(1, h, 1).expand(n, h, w)

    def tap(yy, xx):
        ok = (xx >= 0) & (xx <= w - 1) & (yy >= 0) & (yy <= h - 1)
        idx = (yy.clamp(0, h - 1) * w + xx.clamp(0, w - 1)).view(n, 1, h * w).expand(n, c, h * w)
        v = torch.gather(fea_r.reshape(n, c, h * w), 2, idx).view(n, c, h, w)
        return v * ok.unsqueeze(1).to(v.dtype)

    nw = (fs * fe).unsqueeze(1)
    ne = (fs * fw).unsqueeze(1)
    sw = (fn * fe).unsqueeze(1)
    se = (fn * fw).unsqueeze(1)
    def fma(a, b, acc):
        return (a.double() * b.double() + acc.double()).float()

    out = tap(y0, x0) * nw
    out = fma(tap(y0, x0 + 1), ne, out)
    out = fma(tap(y0 + 1, x0), sw, out)
    out = fma(tap(y0 + 1, x0 + 1), se, out)
    return out


# --------------------------------------------------------------------------------------
# a7  TileWarping.forward                    (model/stereo/hitnet/propagation.py:61-86)
# --------------------------------------------------------------------------------------
def l1_over_channels(x):
    """torch.norm(x, 1, 1): sequential fp32 channel sum of |x| (CPU kernel order)."""
    a = x.abs()
    acc = a[:, 0].clone()
    for ch in range(1, a.shape[1]):
        acc = acc + a[:, ch]
    return acc.unsqueeze(1)


def tile_warp_cost(plane, fea_l, fea_r, direct=False):
    """[N,3,h,w] planes -> [N,48,h,w]: for k in -1,0,+1 the per-pixel L1 matching cost of the
    4x4 slanted tile at disparity d+k, pixel-unshuffled (channel = k_index*16 + yo*4 + xo)."""
    w_fn = warp_right_direct if direct else warp_right
    out = []
    for k in (-1, 0, 1):
        d = plane_expand(plane[:, 0:1] + k, plane[:, 1:2], plane[:, 2:3], 4)
        cost = l1_over_channels(fea_l - w_fn(fea_r, d))
        out.append(F.pixel_unshuffle(cost, 4))
    return torch.cat(out, 1)


# --------------------------------------------------------------------------------------
# a8  TileUpdate0 / TileUpdate               (model/stereo/hitnet/propagation.py:156-172, 206-248)
# --------------------------------------------------------------------------------------
def _resblock(sd, name, x, dilation=1):
    """Sequential(BasicBlock, LeakyReLU): propagation.py:103-121 (padding = dilation)."""
    y = _lrelu(_conv(sd, name + ".0.conv1.0.0", x, padding=dilation, dilation=dilation))
    y = _conv(sd, name + ".0.conv2.0", y, padding=dilation, dilation=dilation)
    return _lrelu(y + x)


def _augment(sd, fea_l, fea_r, hyp, fea_norm, direct):
    cv = tile_warp_cost(hyp[:, :3], fea_l, fea_r, direct)
    return _lrelu(_conv(sd, "decrease.0", torch.cat([fea_norm, cv], 1)))


def tile_update0(sd, fea_l, fea_r, cur, direct=False):
    fea_norm = F.pixel_unshuffle(l1_over_channels(fea_l), 4)
    aug = torch.cat([cur, _augment(sd, fea_l, fea_r, cur, fea_norm, direct)], 1)
    u = _lrelu(_conv(sd, "conv0.0", aug))
    u = _resblock(sd, "resblock0", u)
    u = _resblock(sd, "resblock1", u)
    u = _conv(sd, "lastconv", u, padding=1)
    ref = cur + u
    return torch.cat([F.relu(ref[:, :1]), ref[:, 1:]], 1)


def tile_update(sd, fea_l, fea_r, cur, prev, direct=False, return_aux=False):
    fea_norm = F.pixel_unshuffle(l1_over_channels(fea_l), 4)
    cur_cv = _augment(sd, fea_l, fea_r, cur, fea_norm, direct)
    up_prev = plane_upsample(prev, 2, 2)
    prev_cv = _augment(sd, fea_l, fea_r, up_prev, fea_norm, direct)
    aug = torch.cat((cur, cur_cv, up_prev, prev_cv), 1)
    u = _lrelu(_conv(sd, "conv0.0", aug))
    u = _resblock(sd, "resblock0", u)
    u = _resblock(sd, "resblock1", u)
    u = _conv(sd, "lastconv", u, padding=1)
    refined, new_cur, new_prev, sel = hyp_select(u, cur, up_prev)
    if return_aux:
        return refined, dict(update=u, aug=aug, select=sel, cur=new_cur, prev=new_prev)
    return refined


def hyp_select(update, cur, up_prev):
    """propagation.py:225-248.  update channels: [conf_prev, conf_cur, dprev(16), dcur(16)];
    arg-max over the two confidences, first index on ties (=> previous)."""
    conf = update[:, :2]
    sel = torch.max(conf, dim=1, keepdim=True)[1].float()
    new_cur = cur + update[:, 18:34]
    new_cur = torch.cat([F.relu(new_cur[:, :1]), new_cur[:, 1:]], 1)
    new_prev = up_prev + update[:, 2:18]
    new_prev = torch.cat([F.relu(new_prev[:, :1]), new_prev[:, 1:]], 1)
    refined = sel * new_cur + (1 - sel) * new_prev
    return refined, new_cur, new_prev, sel


# --------------------------------------------------------------------------------------
# a9  PostTileUpdate / FinalTileUpdate       (model/stereo/hitnet/propagation.py:282-290, 325-333)
# --------------------------------------------------------------------------------------
def post_tile_update(sd, fea_l, prev, n_blocks=4, final=False):
    x = torch.cat([fea_l, prev], 1)
    x = _lrelu(_conv(sd, "conv1.0", x))
    x = _lrelu(_conv(sd, "conv1.2", x, padding=1))
    for i in range(n_blocks):
        x = _resblock(sd, f"resblocks.{i}", x, dilation=3 if (i == 1 and not final) else 1)
    x = _conv(sd, "lastconv", x, padding=1)
    if final:
        return F.relu(prev[:, 0:1] + x)
    ref = prev + x
    return torch.cat([F.relu(ref[:, :1]), ref[:, 1:]], 1)


# --------------------------------------------------------------------------------------
# a10 TilePropagation.forward (eval) / HITNetMF.stereo_matching
#     (model/stereo/hitnet/propagation.py:359-372,453; model/stereo/hitnet/hitnet.py:75-100)
# --------------------------------------------------------------------------------------
def tile_propagation(sd, fea_l, fea_r, hyps, direct=False, return_all=False):
    t16 = tile_update0(_sub(sd, "tile_update0."), fea_l[0], fea_r[0], hyps[0], direct)
    levels = [t16]
    prev = t16
    for k in range(1, 5):
        prev = tile_update(_sub(sd, f"tile_update{k}."), fea_l[k], fea_r[k], hyps[k], prev, direct)
        levels.append(prev)
    r1 = post_tile_update(_sub(sd, "tile_update4_1."), fea_l[2], prev, 4)
    r05 = post_tile_update(_sub(sd, "tile_update5."), fea_l[3], plane_upsample(r1, 1, 2), 4)
    r025 = post_tile_update(_sub(sd, "tile_update6."), fea_l[4], plane_upsample(r05, 1, 2), 2, final=True)
    disp = r025[:, 0:1]
    if return_all:
        return disp, dict(levels=levels, r1=r1, r05=r05)
    return disp


def stereo_matching(sd, left, right, max_disp, direct=False, return_all=False, reference_form=False):
    """Eval-mode HITNetMF forward: dict(pred_disp, left_feat, right_feat, left_img).
    ``reference_form``: build the cost volumes with the reference's own op sequence (timing arm)."""
    with torch.no_grad():
        bsd = _sub(sd, "backbone.")
        fl = backbone(bsd, left)
        fr = backbone(bsd, right)
        isd = _sub(sd, "tile_init.")
        tiles = tile_features(isd, fl, fr)
        hyps = tile_hypotheses(isd, tiles, fl, max_disp, reference_form=reference_form)
        res = tile_propagation(_sub(sd, "tile_update."), fl, fr, hyps, direct, return_all)
    disp, extra = res if return_all else (res, None)
    out = dict(pred_disp=disp, left_feat=fl[2], right_feat=fr[2], left_img=left)
    if return_all:
        out.update(fea_l=fl, fea_r=fr, tiles=tiles, hyps=hyps, **extra)
    return out


def stereo_matching_given_argmin(sd, left, right, max_disp, other_disp, rel_tol=1e-5, direct=True, abs_tol=1e-4):
    """End-to-end parity MODULO CERTIFIED NEAR-TIES of the arg-min initialisation.

    The reference's cost volumes contain near-ties at fp32 resolution (e.g. 2.6e-9 absolute / 1.5e-5 relative between
    the two best disparities of a tile in tests/golden/hitnet_s_128x192_d64): an implementation whose feature maps
    differ from the reference's by one ulp may legitimately pick the other disparity, and that one discrete choice
    changes the propagated disparity of a whole image region.  ``other_disp`` is the other implementation's arg-min
    pyramid (5 tensors [N,1,h,w] or [N,h,w], coarse->fine).  Every tile where it differs from this oracle's arg-min is
    checked against THIS oracle's cost volume: the cost at the other disparity must lie within ``max(rel_tol * min,
    abs_tol)`` of the minimum (otherwise the difference is a real error and is reported in ``uncertified``).  The
    absolute bound is what the measured rounding of the tile features explains: a cost is a sum of 16 |L - R| terms, so
    with a feature error e it moves by at most 32 e and two costs can swap order only within 64 e; on the B200 the
    tile features differ from this oracle's by at most 1.2e-6 at 576x960 (profiles/parity_r02.json) -> 7.4e-5.  (A purely
    relative bound is the wrong measure: good matches have a minimum cost near zero.)  The certified choices are adopted and
    the propagation is re-run, so the returned ``pred_disp`` is what the reference computes for those choices.
    Returns dict(pred_disp, flips, uncertified)."""
    with torch.no_grad():
        bsd = _sub(sd, "backbone.")
        fl, fr = backbone(bsd, left), backbone(bsd, right)
        isd = _sub(sd, "tile_init.")
        tiles = tile_features(isd, fl, fr)
        hyps, cvs = tile_hypotheses(isd, tiles, fl, max_disp, return_cv=True)
        flips = uncertified = 0
        for k in range(5):
            od = other_disp[k].reshape(hyps[k][:, 0].shape).to(hyps[k].dtype)
            diff = od != hyps[k][:, 0]
            if diff.any():
                cmin = cvs[k].min(1)[0]
                cother = cvs[k].gather(1, od.long().clamp(0, cvs[k].shape[1] - 1).unsqueeze(1)).squeeze(1)
                near = (cother - cmin) <= (rel_tol * cmin.abs()).clamp(min=abs_tol)
                flips += int(diff.sum())
                uncertified += int((diff & ~near).sum())
                take = diff & near
                hyps[k] = hyps[k].clone()
                hyps[k][:, 0] = torch.where(take, od, hyps[k][:, 0])
        disp = tile_propagation(_sub(sd, "tile_update."), fl, fr, hyps, direct)
    return dict(pred_disp=disp, left_feat=fl[2], right_feat=fr[2], left_img=left, flips=flips, uncertified=uncertified)


# --------------------------------------------------------------------------------------
# parameters and synthetic inputs (Sets U / S / G)
# --------------------------------------------------------------------------------------
def _conv_param(g, cout, cin, kh, kw, transpose=False):
    """torch's default Conv2d init (kaiming_uniform a=sqrt(5) -> U(-1/sqrt(fan_in), ..))."""
    fan_in = (cout if transpose else cin) * kh * kw
    bound = 1.0 / math.sqrt(fan_in)
    shape = (cin, cout, kh, kw) if transpose else (cout, cin, kh, kw)
    w = (torch.rand(shape, generator=g) * 2 - 1) * bound
    b = (torch.rand(cout, generator=g) * 2 - 1) * bound
    return w, b


def hitnet_param_shapes():
    """(name, cout, cin, kh, kw, transpose) for all 106 conv layers = 212 tensors
    (SURVEY.md Appendix B)."""
    L = []
    add = lambda n, co, ci, k, t=False: L.append((n, co, ci, k, k, t))
    add("backbone.conv1.0", 16, 3, 3)
    for name, ci, co in (("down1", 16, 16), ("down2", 16, 24), ("down3", 24, 24)):
        add(f"backbone.{name}.0", co, ci, 4)
        add(f"backbone.{name}.2", co, co, 3)
    add("backbone.down4.0.0", 32, 24, 4)
    add("backbone.down4.0.2", 32, 32, 3)
    add("backbone.down4.1", 32, 32, 3)
    add("backbone.down4.3", 32, 32, 3)
    for name, ci, co in (("up4", 32, 24), ("up3", 24, 24), ("up2", 24, 16), ("up1", 16, 16)):
        add(f"backbone.{name}.0", co, ci, 2, True)
    for name, c in (("merge4", 24), ("merge3", 24), ("merge2", 16), ("merge1", 16)):
        add(f"backbone.{name}.0", c, 2 * c, 1)
        add(f"backbone.{name}.2", c, c, 3)
        add(f"backbone.{name}.4", c, c, 3)
    for name, ci in (("1x", 16), ("2x", 16), ("4x", 24), ("8x", 24), ("16x", 32)):
        add(f"tile_init.tile_conv{name}.0", 16, ci, 4)
        add(f"tile_init.tile_conv{name}.2", 16, 16, 1)
    for name, ci in (("16x", 17), ("8x", 17), ("4x", 33), ("2x", 25), ("1x", 25)):
        add(f"tile_init.tile_fea_dscrpt{name}.0", 13, ci, 1)
    for k in range(5):
        p = f"tile_update.tile_update{k}"
        add(f"{p}.decrease.0", 16, 64, 1)
        add(f"{p}.conv0.0", 32, 32 if k == 0 else 64, 1)
        for r in (0, 1):
            add(f"{p}.resblock{r}.0.conv1.0.0", 32, 32, 3)
            add(f"{p}.resblock{r}.0.conv2.0", 32, 32, 3)
        add(f"{p}.lastconv", 16 if k == 0 else 34, 32, 3)
    for name, cin, hid, cout, nb in (("tile_update4_1", 40, 32, 16, 4), ("tile_update5", 32, 32, 16, 4),
                                     ("tile_update6", 32, 16, 3, 2)):
        p = f"tile_update.{name}"
        add(f"{p}.conv1.0", hid, cin, 1)
        add(f"{p}.conv1.2", hid, hid, 3)
        for r in range(nb):
            add(f"{p}.resblocks.{r}.0.conv1.0.0", hid, hid, 3)
            add(f"{p}.resblocks.{r}.0.conv2.0", hid, hid, 3)
        add(f"{p}.lastconv", cout, hid, 3)
    return L


def random_hitnet_params(seed=0):
    """Random-init HITNetMF state_dict with the reference's names/shapes and torch's
    default init distribution (values differ from ``torch.manual_seed`` module init)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, co, ci, kh, kw, t in hitnet_param_shapes():
        w, b = _conv_param(g, co, ci, kh, kw, t)
        sd[name + ".weight"] = w
        sd[name + ".bias"] = b
    return sd


def params_digest(sd):
    """sha1 over the raw fp32 bytes of a state_dict (key-sorted): exact, order-independent of threads."""
    import hashlib
    hsh = hashlib.sha1()
    for k in sorted(sd):
        hsh.update(sd[k].detach().contiguous().numpy().tobytes())
    return hsh.digest()


def synth_pair(n, h, w, max_disp, seed=1234, kind="S"):
    """Synthetic stereo pair [N,3,h,w] x2 (h, w already multiples of 64).

    U: uniform [0,1), right == left (benchmark_speed.py:40-42; degenerate, timing only).
    S: band-limited texture in ImageNet-normalised range, right = left shifted by a smooth
       disparity field in [0, 0.8*max_disp)  (parity set).
    G: iid normal (kernel-level tests)."""
    g = torch.Generator().manual_seed(seed)
    if kind == "U":
        left = torch.rand(n, 3, h, w, generator=g)
        return left, left.clone()
    if kind == "G":
        return torch.randn(n, 3, h, w, generator=g), torch.randn(n, 3, h, w, generator=g)
    tex = torch.rand(n, 3, h, w + max_disp, generator=g)
    tex = F.avg_pool2d(F.pad(tex, (2, 2, 2, 2), mode="reflect"), 5, stride=1)
    tex = (tex - tex.mean()) / tex.std() * 1.1 + 0.2
    yy = torch.linspace(0, 1, h).view(1, h, 1)
    xx = torch.linspace(0, 1, w).view(1, 1, w)
    ph = torch.rand(n, 1, 1, generator=g) * 6.28
    d = 0.8 * max_disp * (0.5 + 0.25 * torch.sin(3.1 * xx + ph) * torch.cos(2.3 * yy) + 0.2 * yy)
    d = d.clamp(0, 0.8 * max_disp)
    base = torch.arange(w, dtype=torch.float32).view(1, 1, w) + max_disp
    left = _sample_x(tex, base.expand(n, h, w))
    right = _sample_x(tex, base + d)      # right[x] = left[x + d]  <=>  left[x] = right[x - d]
    return left, right


def _sample_x(tex, xs):
    n, c, h, wt = tex.shape
    x0 = xs.floor().clamp(0, wt - 2)
    f = (xs - x0).unsqueeze(1)
    i0 = x0.long().unsqueeze(1).expand(n, c, h, xs.shape[2])
    return torch.gather(tex, 3, i0) * (1 - f) + torch.gather(tex, 3, i0 + 1) * f
