"""Sequence inference driver.

Replaces the host orchestration of apis/inference.py:16-154 (`single_gpu_inference` / `multi_gpu_inference`: iterate
sequences, call the model, write results, gather statistics) and the stereo part of the per-frame loop of
model/codd.py:290-398 for stereo-only models:

  * frames cross PCIe as uint8 (N1: `codd_stage_images_u8` normalises, pads and transposes them on the GPU);
  * per input shape the work of a frame batch — staging of both views, the stereo forward, the crop to the
    image size — is captured ONCE into a CUDA graph and replayed (≈115 kernel launches per batch otherwise);
  * `n_streams` serving slots are used alternately, each with its own pinned host buffers, static device buffers and
    graph, so the H2D copy of batch i+1 and the D2H read of batch i-1 overlap the kernels of batch i;
  * sequences are sharded over ranks as the reference's DistributedSampler does (index i -> rank i % world), no
    data-path collective;
  * results are written as `<name>.disp.pred.npz` (codd.py:596-599) and, with ground truth, accumulated by
    `SequenceMetrics` (N3) without per-frame host synchronisation.
Models with motion / fusion stages keep temporal state and data-dependent control flow between frames: they run
through `model(return_loss=False, ...)` eagerly (`run_model`), one sequence at a time.
"""
import os

import numpy as np
import torch

from . import ops
from .io import write_disp_npz


def shard_indices(n_items, rank=0, world_size=1):
    """Indices of the sequences rank `rank` processes (DistributedSampler(shuffle=False) without padding)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank must be in [0, world_size)")
    return list(range(rank, n_items, world_size))


class _Slot:
    """One serving slot for one input shape: pinned staging buffers, static device buffers, the captured graph."""

    def __init__(self, runner, n, h, w):
        dev = runner.device
        self.shape = (n, h, w)
        self.stream = torch.cuda.Stream(device=dev)
        self.left_h = torch.empty((n, h, w, 3), dtype=torch.uint8).pin_memory()
        self.right_h = torch.empty((n, h, w, 3), dtype=torch.uint8).pin_memory()
        self.left_d = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
        self.right_d = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
        self.out_d = torch.empty((n, 1, h, w), dtype=torch.float32, device=dev)
        self.out_h = torch.empty((n, 1, h, w), dtype=torch.float32).pin_memory()
        self.done = torch.cuda.Event()
        self.graph = None
        self.busy = False
        with torch.cuda.stream(self.stream), torch.no_grad():
            self.left_d.zero_()
            self.right_d.zero_()
            for _ in range(2):                       # warm-up: lazy initialisation (weight packing, smem attributes)
                self._body(runner)
            if runner.use_graph:
                self.stream.synchronize()
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=self.stream):
                    self._body(runner)
        self.stream.synchronize()

    def _body(self, runner):
        n, h, w = self.shape
        left = ops.stage_images_u8(self.left_d, **runner.norm)
        right = ops.stage_images_u8(self.right_d, **runner.norm)
        disp = runner.stereo.stereo_matching(left, right)["pred_disp"]
        self.out_d.copy_(disp[:, :, :h, :w])         # crop to the image size (codd.py:321)

    def launch(self, runner, left_u8, right_u8):
        src_l, src_r = torch.as_tensor(left_u8), torch.as_tensor(right_u8)
        if not src_l.is_pinned():                     # pageable frames are staged through this slot's pinned buffers
            self.left_h.copy_(src_l)
            src_l = self.left_h
        if not src_r.is_pinned():
            self.right_h.copy_(src_r)
            src_r = self.right_h
        with torch.cuda.stream(self.stream), torch.no_grad():
            self.left_d.copy_(src_l, non_blocking=True)
            self.right_d.copy_(src_r, non_blocking=True)
            if self.graph is not None:
                self.graph.replay()
            else:
                self._body(runner)
            self.out_h.copy_(self.out_d, non_blocking=True)
            self.done.record(self.stream)
        self.busy = True

    def wait(self):
        self.done.synchronize()
        self.busy = False
        return self.out_h


class StereoSequenceRunner:
    def __init__(self, model, device="cuda", use_graph=True, n_streams=2, img_norm=None):
        self.stereo = getattr(model, "stereo", model)
        self.stereo.eval()
        self.device = torch.device(device)
        self.use_graph = use_graph
        self.n_streams = n_streams
        cfg = dict(ops.IMG_NORM if img_norm is None else img_norm)
        self.norm = dict(mean=cfg["mean"], std=cfg["std"], to_rgb=cfg.get("to_rgb", True))
        self._slots = {}
        self._turn = {}
        self.graphs_captured = 0

    def _slot(self, n, h, w):
        key = (n, h, w)
        if key not in self._slots:
            self._slots[key] = [_Slot(self, n, h, w) for _ in range(self.n_streams)]
            self._turn[key] = 0
            self.graphs_captured += self.n_streams if self.use_graph else 0
        i = self._turn[key]
        self._turn[key] = (i + 1) % self.n_streams
        return self._slots[key][i]

    def infer_batches(self, batches):
        """batches: iterable of (left_u8, right_u8) uint8 arrays [N,H,W,3] (BGR as cv2 loads them).  Yields the
        disparity maps [N,1,H,W] (numpy, float32) in order; up to n_streams batches are in flight."""
        pending = []
        for left, right in batches:
            n, h, w, _ = left.shape
            slot = self._slot(n, h, w)
            if slot.busy:                              # oldest work on this slot: hand its result out first
                while pending:
                    s = pending.pop(0)
                    yield s.wait().numpy().copy()
                    if s is slot:
                        break
            slot.launch(self, left, right)
            pending.append(slot)
        for s in pending:
            yield s.wait().numpy().copy()

    def run_sequences(self, sequences, out_dir=None, rank=0, world_size=1, metrics=None, batch=1, stats=None):
        """sequences: list of dicts {"name": str, "left": [T,H,W,3] uint8, "right": [T,H,W,3] uint8, optional
        "gt_disp": [T,1,H,W] float32, "gt_flow": [T,2,H,W]}.  This rank processes sequences rank::world_size, `batch`
        frames per launch (stereo frames are independent).  Returns {name: [T,1,H,W] disparity}.

        Output layout follows the reference's show_result (codd.py:596-599, apis/inference.py:50-67): ONE file per
        sequence, `<out_dir>/<name>.disp.pred.npz`, key "disp", shape [1,T,H,W] (batch of one sequence, MF frames).

        Evaluation follows the reference per SEQUENCE (reset_inference_state codd.py:400-433 + RunningStats of
        apis/inference.py:47,72-76): the meters are reset at every sequence start, each sequence yields one row of
        metrics, and `stats` (a SequenceStats, created on demand and returned as results["__stats__"]) holds the
        mean / std over sequences.  `metrics` may be passed to reuse a SequenceMetrics buffer; it is re-sized to the
        sequence length, so long datasets never hit a frame cap."""
        from .metrics import SequenceMetrics, SequenceStats
        results = {}
        for si in shard_indices(len(sequences), rank, world_size):
            seq = sequences[si]
            t_total = seq["left"].shape[0]
            chunks = [(t0, min(t0 + batch, t_total)) for t0 in range(0, t_total, batch)]
            outs = list(self.infer_batches((seq["left"][a:b], seq["right"][a:b]) for a, b in chunks))
            disp = np.concatenate(outs, 0)
            results[seq["name"]] = disp
            if out_dir is not None:
                write_disp_npz(os.path.join(out_dir, seq["name"] + ".png"), disp[None, :, 0])
            if "gt_disp" in seq and (metrics is not None or stats is not None):
                if stats is None:
                    stats = SequenceStats()
                if metrics is None or metrics.acc.shape[0] < t_total:
                    rng = metrics.disp_range if metrics is not None else seq.get("disp_range", (1.0, 192.0))
                    metrics = SequenceMetrics(rng, max_frames=t_total, device=self.device)
                metrics.reset()                        # every meter restarts with the sequence (codd.py:400-433)
                for t in range(t_total):
                    pred = torch.from_numpy(disp[t:t + 1]).to(self.device)
                    gt = torch.as_tensor(seq["gt_disp"][t:t + 1]).to(self.device)
                    flow = None if "gt_flow" not in seq else torch.as_tensor(seq["gt_flow"][t:t + 1]).to(self.device)
                    metrics.update(pred, gt, gt_flow=flow)
                stats.push(seq["name"], metrics.collect())
        if stats is not None:
            results["__stats__"] = stats
        return results


def run_model(model, img, r_img, img_metas, evaluate=False):
    """Full CODD (stereo + motion + fusion): the reference-facing call, eagerly (see module docstring)."""
    with torch.no_grad():
        return model(return_loss=False, rescale=True, evaluate=evaluate, img=img, img_metas=img_metas, r_img=r_img)
