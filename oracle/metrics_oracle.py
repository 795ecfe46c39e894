"""CPU restatement of the reference's per-frame evaluation arithmetic.

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing on the product path).  numpy, float32
arithmetic step by step where the reference computes in float32, float64 only for the final sums.

Follows
  utils/misc.py:12-36      compute_valid_mask
  utils/warp.py:7-40,69-92 normalize_coords / meshgrid / flow_warp (mode="nearest", padding_mode="zeros";
                           F.grid_sample with align_corners=True restated: unnormalise ((c+1)/2)*(size-1),
                           nearbyint, zero outside the image)
  utils/metric.py:9-54     epe_metric, t_epe_metric, thres_metric
  model/codd.py:435-517    calc_metric: disparity block (:462-474) and temporal block (:476-515)
  model/codd.py:519-575    the scene-flow block (induced_flow of projective_ops.py:11-68; see the note further down)
Pinned against those functions imported from /root/reference in tests/test_metrics_oracle.py (disparity and
temporal blocks); the scene-flow block's `Ts * X0` is lietorch's and therefore unpinned.
"""
import numpy as np

BF_DEFAULT = np.float32(1050 * 0.2)   # utils/misc.py:7
F32 = np.float32


def valid_mask(gt_disp, disp_range, seg=None, flow_prev=None, disp_change=None):
    """utils/misc.py:26-33.  gt_disp [N,1,H,W]; seg [N,1,H,W]; flow_prev [N,2,H,W] -> bool [N,1,H,W]."""
    gt = np.asarray(gt_disp, F32)
    m = (gt > F32(disp_range[0])) & (gt < F32(disp_range[1]))
    if seg is not None:
        m &= np.asarray(seg) > 0
    if flow_prev is not None:
        f = np.asarray(flow_prev, F32)
        mag = np.sqrt((f[:, 0:1] * f[:, 0:1] + f[:, 1:2] * f[:, 1:2]).astype(F32)).astype(F32)
        m &= mag < BF_DEFAULT
    if disp_change is not None:
        m &= np.abs(np.asarray(disp_change, F32)) < BF_DEFAULT
    return m


def _sample_index(base, flow, size):
    """grid + flow -> normalize_coords -> grid_sample unnormalise -> nearbyint, all in float32."""
    s = (base + flow).astype(F32)
    n = (F32(2) * (s / F32(size - 1)).astype(F32)).astype(F32) - F32(1)          # warp.py:14-15
    u = (((n + F32(1)).astype(F32) / F32(2)).astype(F32) * F32(size - 1)).astype(F32)
    return np.rint(u)                                                           # round half to even


def flow_warp_nearest(img, flow):
    """utils/warp.py:69-92 with mode='nearest', padding_mode='zeros'.  img [N,C,H,W], flow [N,2,H,W]
    -> warped [N,C,H,W] float32, valid [N,C,H,W] bool."""
    img = np.asarray(img, F32)
    flow = np.asarray(flow, F32)
    n, c, h, w = img.shape
    xs = np.arange(w, dtype=F32)[None, None, :]
    ys = np.arange(h, dtype=F32)[None, :, None]
    ix = _sample_index(np.broadcast_to(xs, (n, h, w)), flow[:, 0], w)
    iy = _sample_index(np.broadcast_to(ys, (n, h, w)), flow[:, 1], h)
    inb = (ix >= 0) & (ix <= w - 1) & (iy >= 0) & (iy <= h - 1)
    xi = np.clip(ix, 0, w - 1).astype(np.int64)
    yi = np.clip(iy, 0, h - 1).astype(np.int64)
    out = np.zeros_like(img)
    for b in range(n):
        g = img[b][:, yi[b], xi[b]]
        out[b] = np.where(inb[b][None], g, F32(0))
    valid = np.broadcast_to(inb[:, None], img.shape).copy()
    return out, valid


def disp_metrics(pred, gt, mask):
    """codd.py:462-474.  Returns dict(n, epe, th3); epe / th3 are None when no pixel is valid."""
    pred, gt = np.asarray(pred, F32), np.asarray(gt, F32)
    n = int(mask.sum())
    if n == 0:
        return dict(n=0, epe=None, th3=None)
    e = np.abs((pred[mask] - gt[mask]).astype(F32))
    return dict(n=n, epe=float(e.astype(np.float64).sum() / n), th3=float((e > F32(3.0)).sum() / n))


def temporal_metrics(flow, gt, pred, seg, gt_prev, pred_prev, mask_prev, disp_range, gt_disp2_prev=None):
    """codd.py:476-515 for one frame pair.  flow = ground-truth flow of the PREVIOUS frame [N,2,H,W]; gt / pred the
    current frame [N,1,H,W]; *_prev the previous one; mask_prev the previous frame's disparity mask.
    Returns dict(updated, tepe, tepe_rel, th1_tepe_rel, th3_tepe, n, flow_mag)."""
    gt, pred = np.asarray(gt, F32), np.asarray(pred, F32)
    flow = np.asarray(flow, F32)
    if (gt > 0).any():
        mask = valid_mask(gt, disp_range, seg=seg, flow_prev=flow)
    else:   # KITTI: only one frame has disparity, dummy gt of BF/2 (codd.py:486-490)
        mask = valid_mask(np.full_like(gt, BF_DEFAULT / F32(2.0)), disp_range, seg=seg, flow_prev=flow)
    to_warp = np.concatenate([gt, pred, mask.astype(F32)], axis=1)
    warped, valid = flow_warp_nearest(to_warp, flow)
    w_gt, w_pred, w_mask = warped[:, 0:1], warped[:, 1:2], warped[:, 2:3]
    # codd.py:499 takes valid.squeeze()[0]: with the batch of 1 the reference evaluates at, channel 0's mask
    mask_curr = valid[:, 0:1] & (w_mask != 0) & mask
    if gt_disp2_prev is not None:
        w_gt = np.asarray(gt_disp2_prev, F32)
        mask_curr = mask_curr & (w_gt > 0)
    mag = np.sqrt((flow[:, 0] * flow[:, 0] + flow[:, 1] * flow[:, 1]).astype(F32)).astype(F32)
    out = dict(updated=False, flow_mag=float(mag.astype(np.float64).mean()))
    if mask_prev.any() and mask_curr.any():
        d_est = (w_pred - np.asarray(pred_prev, F32)).astype(F32)
        d_gt = (w_gt - np.asarray(gt_prev, F32)).astype(F32)
        m = mask_prev & mask_curr
        abs_err = np.abs((d_est - d_gt).astype(F32))[m]
        rel = (abs_err / (np.abs(d_gt[m]) + F32(1e-3)).astype(F32)).astype(F32)
        n = int(m.sum())
        with np.errstate(invalid="ignore", divide="ignore"):
            out.update(updated=True, n=n,
                       tepe=float(abs_err.astype(np.float64).sum() / n) if n else float("nan"),
                       tepe_rel=float(rel.astype(np.float64).sum() / n) if n else float("nan"),
                       th1_tepe_rel=float((rel > F32(1.0)).sum() / n) if n else float("nan"),
                       th3_tepe=float((abs_err > F32(3.0)).sum() / n) if n else float("nan"))
    return out


# ----------------------------------------------------------------------------------------------
# scene-flow block (codd.py:519-575): induced_flow (model/motion/raft3d/projective_ops.py:11-68) restated for a dense
# SE3 field given as (tx, ty, tz, qx, qy, qz, qw).  The lietorch action `Ts * X0` is restated as the quaternion
# rotation + translation (lietorch is absent from this container: that one line is unpinned; the projective ops are
# the ones pinned in oracle/motion_oracle.py).
# ----------------------------------------------------------------------------------------------
P_EPS = F32(1e-5)   # projective_ops.py:8


def induced_flow(Ts, depth, intr):
    """Ts [N,H,W,7], depth [N,H,W], intr [N,4] -> flow2d [N,H,W,3] = project(T*X0) - project(X0) (float32)."""
    Ts, depth, intr = np.asarray(Ts, F32), np.asarray(depth, F32), np.asarray(intr, F32)
    n, h, w = depth.shape
    fx, fy, cx, cy = [intr[:, i][:, None, None] for i in range(4)]
    x = np.arange(w, dtype=F32)[None, None, :]
    y = np.arange(h, dtype=F32)[None, :, None]
    X0 = np.stack([depth * ((x - cx) / fx), depth * ((y - cy) / fy), depth], -1).astype(F32)
    qv, qw, t = Ts[..., 3:6], Ts[..., 6:7], Ts[..., 0:3]
    uv = (F32(2) * np.cross(qv, X0)).astype(F32)
    X1 = (X0 + qw * uv + np.cross(qv, uv) + t).astype(F32)

    def project(X):
        Z = X[..., 2] + P_EPS
        return np.stack([fx * (X[..., 0] / Z) + cx, fy * (X[..., 1] / Z) + cy, F32(1) / Z], -1).astype(F32)

    return (project(X1) - project(X0)).astype(F32)


def sceneflow_metrics(Ts, pred_prev, intr, flow, gt_disp_change, gt_disp_prev, disp_range, seg=None, flow_occ=None):
    """codd.py:519-575 for one frame pair.  Returns dict(n, sum_sf, sum_of, n1_sf, n1_of)."""
    flow = np.asarray(flow, F32)
    mask = valid_mask(gt_disp_prev, disp_range, seg=seg, flow_prev=flow, disp_change=gt_disp_change)
    if flow_occ is not None:
        mask = mask & ~np.asarray(flow_occ, bool)
    with np.errstate(divide="ignore", invalid="ignore"):
        depth1 = np.clip((BF_DEFAULT / np.asarray(pred_prev, F32)).astype(F32), F32(0), BF_DEFAULT)[:, 0]
    est = induced_flow(Ts, depth1, intr)
    est[..., 2] = est[..., 2] * BF_DEFAULT
    gt3 = np.concatenate([flow.transpose(0, 2, 3, 1), np.asarray(gt_disp_change, F32).transpose(0, 2, 3, 1)], -1)
    d2 = ((est - gt3).astype(F32) ** 2).astype(F32)
    sf = np.sqrt(d2.sum(-1, dtype=F32))
    of = np.sqrt(d2[..., :2].sum(-1, dtype=F32))
    m = mask[:, 0]
    return dict(n=int(m.sum()), sum_sf=float(sf[m].astype(np.float64).sum()), sum_of=float(of[m].astype(np.float64).sum()),
                n1_sf=int((sf[m] < F32(1.0)).sum()), n1_of=int((of[m] < F32(1.0)).sum()))


def gt_disp_change(flow_occ_prev, gt_prev, gt_curr, flow):
    """utils/misc.py:39-59 compute_gt_disp_change -> (change, warped gt), float32 [N,1,H,W]."""
    warped, valid = flow_warp_nearest(gt_curr, flow)
    change = (warped - np.asarray(gt_prev, F32)).astype(F32)
    change[~valid] = BF_DEFAULT
    if flow_occ_prev is not None:
        change[np.asarray(flow_occ_prev, bool)] = BF_DEFAULT
    return change, warped
