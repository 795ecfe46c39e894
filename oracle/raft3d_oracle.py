"""CPU oracle for the RAFT3D networks and the Motion module:
BasicEncoder, HRNet + ResizeConcatConv, ConvGRU / BasicUpdateBlock, the RAFT3D iteration loop and
Motion.forward, as plain functional torch over a ``state_dict``.

TEST INFRASTRUCTURE ONLY (rules in oracle/hitnet_oracle.py): only tests/, __graft_entry__.smoke() and
bench.py's CPU legs import this.

PARITY STATUS
  * PINNED against the unmodified reference files (importable through oracle/_shim; see
    tests/test_oracle_vs_reference.py): ``basic_encoder`` (blocks/extractor.py), ``update_block``
    (raft3d.py:43-106 + blocks/gru.py), ``resize_concat`` (raft3d.py:109-137).
  * **PARITY UNPINNED**: ``hrnet`` restates mmseg's HRNet (mmsegmentation 0.x, unpinned, README.md:42 — not
    vendored in /root/reference and not installable here) from its published structure; ``raft3d_forward`` /
    ``motion_forward`` compose the pinned pieces with the unpinned lietorch / lietorch_extras / pytorch3d
    restatements of oracle/motion_oracle.py, following the reference's call sequence line by line
    (raft3d.py:190-280, motion.py:132-209).
"""
import torch
import torch.nn.functional as F

from . import motion_oracle as M

BF_DEFAULT = 1050 * 0.2

HRNET_EXTRA = dict(      # configs/models/codd.py:48-73
    stage1=dict(num_modules=1, num_branches=1, block="BOTTLENECK", num_blocks=(2,), num_channels=(64,)),
    stage2=dict(num_modules=1, num_branches=2, block="BASIC", num_blocks=(2, 2), num_channels=(18, 36)),
    stage3=dict(num_modules=3, num_branches=3, block="BASIC", num_blocks=(2, 2, 2), num_channels=(18, 36, 72)),
    stage4=dict(num_modules=2, num_branches=4, block="BASIC", num_blocks=(2, 2, 2, 2), num_channels=(18, 36, 72, 144)))


def _conv(sd, name, x, stride=1, padding=0, dilation=1):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding, dilation=dilation)


# ----------------------------------------------------------------------------------------------
# BasicEncoder, instance norm   (blocks/extractor.py:9-55,124-199)  [PINNED]
# ----------------------------------------------------------------------------------------------
def _res_block(sd, p, x, stride):
    y = F.relu(F.instance_norm(_conv(sd, p + ".conv1", x, stride, 1)))
    y = F.relu(F.instance_norm(_conv(sd, p + ".conv2", y, 1, 1)))
    if stride != 1:
        x = F.instance_norm(_conv(sd, p + ".downsample.0", x, stride))
    return F.relu(x + y)


def basic_encoder(sd, p, x):
    x = F.relu(F.instance_norm(_conv(sd, p + "conv1", x, 2, 3)))
    for layer, stride in (("layer1", 1), ("layer2", 2), ("layer3", 2)):
        x = _res_block(sd, f"{p}{layer}.0", x, stride)
        x = _res_block(sd, f"{p}{layer}.1", x, 1)
    return _conv(sd, p + "conv2", x)


# ----------------------------------------------------------------------------------------------
# HRNet (mmseg 0.x structure)  [UNPINNED]
# ----------------------------------------------------------------------------------------------
def _bn(sd, name, x, eps=1e-5):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], False, 0.0, eps)


def _basic_block(sd, p, x):
    idt = x
    if p + ".downsample.0.weight" in sd:
        idt = _bn(sd, p + ".downsample.1", _conv(sd, p + ".downsample.0", x))
    y = F.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x, 1, 1)))
    y = _bn(sd, p + ".bn2", _conv(sd, p + ".conv2", y, 1, 1))
    return F.relu(y + idt)


def _bottleneck(sd, p, x):
    idt = x
    if p + ".downsample.0.weight" in sd:
        idt = _bn(sd, p + ".downsample.1", _conv(sd, p + ".downsample.0", x))
    y = F.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x)))
    y = F.relu(_bn(sd, p + ".bn2", _conv(sd, p + ".conv2", y, 1, 1)))
    y = _bn(sd, p + ".bn3", _conv(sd, p + ".conv3", y))
    return F.relu(y + idt)


def _hr_module(sd, p, xs, num_blocks):
    nb = len(xs)
    xs = list(xs)
    for i in range(nb):
        for b in range(num_blocks[i]):
            xs[i] = _basic_block(sd, f"{p}.branches.{i}.{b}", xs[i])
    if nb == 1:
        return xs
    outs = []
    for i in range(nb):
        y = 0
        for j in range(nb):
            q = f"{p}.fuse_layers.{i}.{j}"
            if i == j:
                y = y + xs[j]
            elif j > i:
                t = _bn(sd, q + ".1", _conv(sd, q + ".0", xs[j]))
                y = y + F.interpolate(t, size=xs[i].shape[2:], mode="bilinear", align_corners=False)
            else:
                t = xs[j]
                for k in range(i - j):
                    t = _bn(sd, f"{q}.{k}.1", _conv(sd, f"{q}.{k}.0", t, 2, 1))
                    if k != i - j - 1:
                        t = F.relu(t)
                y = y + t
        outs.append(F.relu(y))
    return outs


def hrnet(sd, p, x, extra=HRNET_EXTRA):
    x = F.relu(_bn(sd, p + "bn1", _conv(sd, p + "conv1", x, 2, 1)))
    x = F.relu(_bn(sd, p + "bn2", _conv(sd, p + "conv2", x, 2, 1)))
    for b in range(extra["stage1"]["num_blocks"][0]):
        x = _bottleneck(sd, f"{p}layer1.{b}", x)
    ys = [x]
    for idx in (2, 3, 4):
        cfg = extra[f"stage{idx}"]
        xs = []
        for i in range(cfg["num_branches"]):
            q = f"{p}transition{idx - 1}.{i}"
            if q + ".0.weight" in sd:                       # same-resolution 3x3 conv
                xs.append(F.relu(_bn(sd, q + ".1", _conv(sd, q + ".0", ys[-1], 1, 1))))
            elif q + ".0.0.weight" in sd:                   # new branch: chain of stride-2 convs
                t, k = ys[-1], 0
                while f"{q}.{k}.0.weight" in sd:
                    t = F.relu(_bn(sd, f"{q}.{k}.1", _conv(sd, f"{q}.{k}.0", t, 2, 1)))
                    k += 1
                xs.append(t)
            else:
                xs.append(ys[i])
        for m in range(cfg["num_modules"]):
            xs = _hr_module(sd, f"{p}stage{idx}.{m}", xs, cfg["num_blocks"])
        ys = xs
    return ys


def resize_concat(sd, p, inputs):
    """raft3d.py:125-137  [PINNED]"""
    ups = [F.interpolate(x, size=inputs[1].shape[2:], mode="bilinear", align_corners=True) for x in inputs]
    return F.relu(_conv(sd, p + "convs.0", torch.cat(ups, 1)))


def context_net(sd, p, image):
    return resize_concat(sd, p + "1.", hrnet(sd, p + "0.", image))


# ----------------------------------------------------------------------------------------------
# update block   (raft3d.py:43-106, blocks/gru.py:10-35)  [PINNED]
# ----------------------------------------------------------------------------------------------
def conv_gru(sd, p, h, *inputs):
    iz = ir = iq = 0
    for inp in inputs:
        a, b, c = inp.split([128, 128, 128], 1)
        iz, ir, iq = iz + a, ir + b, iq + c
    z = torch.sigmoid(_conv(sd, p + "convz1", h, 1, 1) + _conv(sd, p + "convz2", h, 1, 4, 4) + iz)
    r = torch.sigmoid(_conv(sd, p + "convr1", h, 1, 1) + _conv(sd, p + "convr2", h, 1, 4, 4) + ir)
    q = torch.tanh(_conv(sd, p + "convq1", r * h, 1, 1) + _conv(sd, p + "convq2", r * h, 1, 4, 4) + iq)
    return (1 - z) * h + z * q


def _head(sd, p, x):
    return _conv(sd, p + ".2", F.relu(_conv(sd, p + ".0", x, 1, 1)))


def update_block(sd, p, net, inp, corr, flow, twist, dz):
    """Argument names follow the SIGNATURE (raft3d.py:92); the reference's call site passes (flow, dz, twist)."""
    info = torch.cat([flow, 10 * dz, 10 * twist], -1).clamp(-50.0, 50.0).permute(0, 3, 1, 2)
    mot = _conv(sd, p + "flow_enc.2", F.relu(_conv(sd, p + "flow_enc.0", info, 1, 3)))
    cor = F.relu(_conv(sd, p + "corr_enc.0", corr, 1, 1))
    cor = _conv(sd, p + "corr_enc.4", F.relu(_conv(sd, p + "corr_enc.2", cor, 1, 1)))
    net = conv_gru(sd, p + "gru.", net, inp, cor, mot)
    return (net, _head(sd, p + "mask", net), _head(sd, p + "ae", net), _head(sd, p + "delta", net),
            torch.sigmoid(_head(sd, p + "weight", net)))


# ----------------------------------------------------------------------------------------------
# RAFT3D.forward / Motion.forward   (raft3d.py:190-280, motion.py:132-209)
# ----------------------------------------------------------------------------------------------
def raft3d_forward(sd, p, image_curr, depth_prev, depth_curr, intrinsics, state, outputs, iters=12):
    if "memory" not in state:
        state["raft_feat"] = basic_encoder(sd, p + "fnet.", image_curr)
        state["raft_netinp"] = context_net(sd, p + "cnet.", image_curr)
        return
    fmap_prev, net_inp = state["raft_feat"], state["raft_netinp"]
    n, _, ht, wd = image_curr.shape
    Ts = M.se3_identity(n, ht // 8, wd // 8)
    y0, x0 = torch.meshgrid(torch.arange(ht // 8).float(), torch.arange(wd // 8).float(), indexing="ij")
    coords0 = torch.stack([x0, y0], -1)[None].repeat(n, 1, 1, 1)
    fmap_curr = basic_encoder(sd, p + "fnet.", image_curr)
    pyramid = M.all_pairs_correlation(fmap_prev, fmap_curr, 4)
    net, inp = net_inp.split([128, 384], 1)
    net, inp = torch.tanh(net), torch.relu(inp)
    intr8 = intrinsics / 8.0
    depth1_r8, depth2_r8 = depth_prev[:, 3::8, 3::8], depth_curr[:, 3::8, 3::8]
    for _ in range(iters):
        xyz, _ = M.projective_transform(Ts, depth1_r8, intr8)
        coords1, zinv_proj = xyz.split([2, 1], -1)
        zinv = M.depth_sampler(1.0 / depth2_r8, coords1)
        corr = M.corr_lookup(pyramid, coords1.permute(0, 3, 1, 2).contiguous(), radius=3)
        flow = coords1 - coords0
        dz = zinv.unsqueeze(-1) - zinv_proj
        twist = M.se3_log(Ts)
        net, mask, ae, delta, weight = update_block(sd, p + "update_block.", net, inp, corr, flow, dz, twist)
        target = (xyz.permute(0, 3, 1, 2) + delta).contiguous()
        Ts = M.gn_step(Ts, ae, target, weight, depth1_r8, intr8)
    Ts_up = M.upsample_se3(Ts, mask)
    outputs["Ts"] = Ts_up
    outputs["flow2d_est_induced"] = M.induced_flow(Ts_up, depth_prev, intrinsics)[0]
    outputs["weight"] = M.cvx_upsample(weight.permute(0, 2, 3, 1), mask).permute(0, 3, 1, 2)
    state["raft_feat"] = fmap_curr
    state["raft_netinp"] = context_net(sd, p + "cnet.", image_curr)


def motion_forward(sd, p, state, outputs, intrinsics, iters=16, ds=4):
    """motion.py:132-209.  ``intrinsics``: python list [fx, fy, cx, cy] (img_metas[0]["intrinsics"])."""
    img_curr = outputs["left_img"]
    if "memory" not in state:
        raft3d_forward(sd, p + "raft3d.", img_curr, None, None, None, state, outputs)
        return
    B = outputs["pred_disp"].shape[0]
    intr = torch.tensor(intrinsics).unsqueeze(0).expand(B, -1)
    depth_scale = BF_DEFAULT / intr[0, 0]
    img_prev, feat_prev, disp_prev = state["memory"]
    disp_curr = outputs["pred_disp"]
    depth_prev = torch.clip(depth_scale * intr[0, 0] / (disp_prev + 1e-5), max=BF_DEFAULT, min=0).reshape(B, *disp_prev.shape[-2:])
    depth_curr = torch.clip(depth_scale * intr[0, 0] / (disp_curr + 1e-5), max=BF_DEFAULT, min=0).squeeze(1)
    intr = intr.float()
    raft3d_forward(sd, p + "raft3d.", img_curr, depth_prev, depth_curr, intr, state, outputs, iters=iters)
    Ts = outputs["Ts"]
    w = depth_curr.shape[-1]
    to_proj = torch.cat([img_prev, outputs["flow2d_est_induced"].permute(0, 3, 1, 2), outputs["weight"]], 1)
    warped, depth_warp = M.splat_warp(Ts, depth_prev, to_proj, intr, radius=2.0)
    disp_warp = depth_scale * intr[0, 0] / (depth_warp + 1e-5)
    disp_warp[disp_warp > w] = 0.0
    Ts_lr = Ts[:, ds // 2 - 1::ds, ds // 2 - 1::ds]
    depth_lr = depth_prev[:, ds // 2 - 1::ds, ds // 2 - 1::ds]
    feat_warp, _ = M.splat_warp(Ts_lr, depth_lr, feat_prev, intr / ds, radius=4.0)
    if disp_warp.dim() == 3:
        disp_warp = disp_warp.unsqueeze(1)
    state["memory"] = [warped[:, :3], feat_warp, warped[:, 6:], disp_warp, warped[:, 3:6]]


# ----------------------------------------------------------------------------------------------
# random parameters with the reference's names / shapes
# ----------------------------------------------------------------------------------------------
def random_raft3d_params(seed=0):
    """state_dict of a randomly initialised RAFT3D (BatchNorm running statistics randomised so that folding is
    exercised).  Built from the codd_b200 module tree, whose parameter names are the reference's; only the
    container classes (no kernels) are touched, on CPU."""
    from codd_b200.motion.raft3d import RAFT3D
    torch.manual_seed(seed)
    net = RAFT3D(cnet_cfg=dict(type="HRNet", norm_cfg=dict(type="SyncBN", requires_grad=False), norm_eval=True,
                               extra=HRNET_EXTRA))
    g = torch.Generator().manual_seed(seed + 1)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    for k in sd:
        if k.endswith("running_mean"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
        elif k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(sd[k].shape, generator=g)
        elif k.endswith("bn1.weight") or k.endswith("bn2.weight") or k.endswith("bn3.weight") or k.endswith(".1.weight"):
            if sd[k].dim() == 1:
                sd[k] = 0.5 + torch.rand(sd[k].shape, generator=g)
        elif k.endswith("bias") and sd[k].dim() == 1 and "cnet.0" in k:
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
    return sd
