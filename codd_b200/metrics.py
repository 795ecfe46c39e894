"""On-GPU evaluation state for sequence inference.

Mirrors what ``ConsistentOnlineDynamicDepth.calc_metric`` / ``reset_inference_state`` keep in
``inference_state`` (model/codd.py:400-517): per frame the EPE / 3-px meters, from the second frame on the temporal
EPE meters and the flow-magnitude meter.  The reference calls ``.item()`` several times per frame (a host
synchronisation each); here every frame adds one row of float64 sums on the device (two kernel launches) and
``collect()`` reads them back once and applies the reference's AverageMeter semantics (mean over the frames that
updated a meter, utils/running_stats.py) and the running sums of the scene-flow block (codd.py:519-575).
"""
import numpy as np
import torch

from . import ops

ROW = 20   # doubles per frame: [0:4] codd_disp_metrics, [4:13] codd_temporal_metrics, [13:18] codd_sceneflow_metrics


class SequenceMetrics:
    def __init__(self, disp_range, max_frames=4096, device="cuda"):
        self.disp_range = (float(disp_range[0]), float(disp_range[1]))
        self.acc = torch.zeros((max_frames, ROW), dtype=torch.float64, device=device)
        self.frames = 0
        self._prev = None   # (gt, pred, mask, flow) of the previous frame

    def reset(self):
        self.acc.zero_()
        self.frames = 0
        self._prev = None

    def update(self, pred_disp, gt_disp, gt_flow=None, seg=None, gt_disp2=None, Ts=None, intrinsics=None,
               gt_disp_change=None, gt_flow_occ_prev=None):
        """One frame (codd.py:435-517).  pred_disp: the network output [N,1,Hp,Wp] (padded is fine, it is cropped to
        gt's size); gt_disp [N,1,H,W]; gt_flow [N,2,H,W] = this frame's ground-truth flow to the NEXT frame (kept for
        the next call, as inference_state["gt_flow"][-2]); seg: optional semantic / occlusion mask (> 0 = keep).
        Motion meters (codd.py:519-575), from the second frame on: Ts = the SE3 field [N,H',W',7] estimated between the
        previous and this frame, intrinsics [N,4], gt_disp_change [N,1,H,W] = the disparity change of the previous
        frame's pixels (the reference's gt_disp_change[-2], or [-1] when it was derived from flow, in which case
        gt_flow_occ_prev = gt_flow_occ[-2] removes the occluded pixels)."""
        if self.frames >= self.acc.shape[0]:
            raise RuntimeError("SequenceMetrics: max_frames exceeded")
        n, _, h, w = gt_disp.shape
        row = self.acc[self.frames]
        mask = torch.empty((n, 1, h, w), dtype=torch.uint8, device=gt_disp.device)
        ops.disp_metrics(pred_disp, gt_disp, self.disp_range, row[0:4], seg=seg, mask_out=mask)
        if self._prev is not None and self._prev[3] is not None:
            p_gt, p_pred, p_mask, p_flow, p_gt2 = self._prev
            ops.temporal_metrics(p_flow, gt_disp, pred_disp, p_gt, p_pred, p_mask, self.disp_range, row[4:13], seg=seg,
                                 gt_disp2_prev=p_gt2, gt_pos_count=row[3:4])
            if Ts is not None and gt_disp_change is not None:
                ops.sceneflow_metrics(Ts, p_pred, intrinsics, p_flow, gt_disp_change, p_gt, self.disp_range, row[13:18],
                                      seg=seg, flow_occ=gt_flow_occ_prev)
        self._prev = (gt_disp, pred_disp, mask, gt_flow, gt_disp2)
        self.frames += 1

    def rows(self):
        """The per-frame accumulator rows of this rank as a CPU float64 tensor [frames, ROW] (one D2H copy)."""
        return self.acc[: self.frames].cpu()

    def collect(self, all_ranks=False):
        """One device->host copy; returns the meters of utils/misc.py:62-86.  all_ranks=True: the rows of every rank
        are gathered first (the role of apis/inference.py's collect_results after multi_gpu_inference), so every rank
        returns the statistics of the whole dataset."""
        rows = self.rows()
        if all_ranks:
            rows = gather_rows(rows)
        return summarise(rows.numpy())


class SequenceStats:
    """Mean / standard deviation over SEQUENCES of the per-sequence metric rows: the role of the reference's
    RunningStatsWithBuffer (utils/running_stats.py, apis/inference.py:47,72-76,140-153): one `collect_metric` row per
    sequence, `mean` / `std` over the rows, ranks merged by gathering the rows."""

    def __init__(self):
        self.names, self.rows = [], []

    def push(self, name, metrics):
        self.names.append(name)
        self.rows.append(dict(metrics))

    @property
    def n(self):
        return len(self.rows)

    def _table(self):
        keys = sorted(self.rows[0]) if self.rows else []
        return keys, np.array([[r[k] for k in keys] for r in self.rows], dtype=np.float64).reshape(len(self.rows), len(keys))

    @property
    def mean(self):
        keys, t = self._table()
        return dict(zip(keys, t.mean(0))) if len(t) else {}

    @property
    def std(self):
        """Population standard deviation over sequences (sqrt(s / n)), as the reference's RunningStats.std."""
        keys, t = self._table()
        return dict(zip(keys, t.std(0))) if len(t) else {}

    def gather(self):
        """All ranks' rows (dist.all_gather_object, as apis/inference.py:147); returns a merged SequenceStats."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return self
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, (self.names, self.rows))
        merged = SequenceStats()
        for names, rows in parts:
            merged.names += names
            merged.rows += rows
        return merged

    def dump(self, path):
        """stats.csv: one line per sequence, then mean and std (RunningStatsWithBuffer.dump)."""
        import csv
        keys, t = self._table()
        with open(path, "w", newline="") as f:
            wr = csv.writer(f)
            wr.writerow(["name"] + keys)
            for name, row in zip(self.names, t):
                wr.writerow([name] + [f"{v:.6g}" for v in row])
            wr.writerow(["mean"] + [f"{self.mean[k]:.6g}" for k in keys])
            wr.writerow(["std"] + [f"{self.std[k]:.6g}" for k in keys])


def gather_rows(rows):
    """All-gather of per-frame accumulator rows [frames_r, ROW] over the default process group (ranks may hold
    different numbers of frames).  Works on NCCL (rows are moved to the current CUDA device) and gloo."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rows
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    world = dist.get_world_size()
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([rows.shape[0]], dtype=torch.int64, device=dev))
    cap = max(int(c.item()) for c in counts)
    pad = torch.zeros((cap, rows.shape[1]), dtype=torch.float64, device=dev)
    pad[: rows.shape[0]] = rows.to(dev)
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: int(c.item())].cpu() for o, c in zip(out, counts)], 0)


def summarise(a):
    """Accumulator rows [frames, ROW] (numpy float64) -> the reference's meters (AverageMeter = mean over the frames
    that updated it, utils/running_stats.py; scene-flow entries are running sums, codd.py:567-575)."""
    out = {}

    def meter(name, num, den, gate):
        vals = [num[i] / den[i] if den[i] else float("nan") for i in range(len(num)) if gate[i]]
        out[name] = float(np.mean(vals)) if vals else 0.0

    has = a[:, 0] > 0
    meter("epe", a[:, 1], a[:, 0], has)
    meter("th3", a[:, 2], a[:, 0], has)
    t = a[:, 4:13]
    upd = (t[:, 5] > 0) & (t[:, 6] > 0)              # mask_prev.any() and mask_curr.any() (codd.py:506)
    meter("tepe", t[:, 1], t[:, 0], upd)
    meter("tepe_rel", t[:, 2], t[:, 0], upd)
    meter("th1_tepe_rel", t[:, 3], t[:, 0], upd)
    meter("th3_tepe", t[:, 4], t[:, 0], upd)
    meter("flow_mag", t[:, 7], t[:, 8], t[:, 8] > 0)
    sf = a[:, 13:18].sum(0)                           # running sums over the sequence (codd.py:567-575, misc.py:76-77)
    out.update(count=float(sf[0]), epe2d_scene_flow=float(sf[1]), epe2d_optical_flow=float(sf[2]),
               **{"1px_scene_flow": float(sf[3]), "1px_optical_flow": float(sf[4])})
    return out
