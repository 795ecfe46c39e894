#!/usr/bin/env python
"""bench.py — stereo frames/s of the CODD HITNetMF hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one HITNetMF stereo-only forward over a batch of 8 synthetic 960x540 pairs
(reflect-padded to 576x960, D=192, random-init weights): BASELINE.json configs[1].  With N GPUs
every rank runs the same per-GPU batch (frames are independent -> weak scaling, no data-path
collective; NCCL only broadcasts the weights once and reduces the timing).

One JSON line on stdout (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = the
same metric through the public model(...) call with pinned HOST buffers, host<->device copies
inside the timed region.  `roofline` describes the dominant kernel of the step (per-launch CUDA
events in an instrumented pass right after the timed region); `cpu_baseline` is the CPU oracle
port timed on this box's host cores on a bounded sample.

--impl reference: times the UNMODIFIED reference HITNetMF (the git-ignored mirror of /root/reference's model/ and
utils/ under baseline/_ref, imported through oracle/_shim because mmcv / mmseg are not installable) on all host
cores; `cpu_baseline.kind` = "reference".  Only if that mirror is missing does it fall back to the bit-identical
oracle port in reference form ("port").  `gpu_eager_baseline` in the main line is the same unmodified reference
model executed by torch-CUDA eager (cuDNN) on the same B200, same batch: the library path the hand-written kernels
have to beat (/root/reference/benchmark_speed.py:44-64).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

H_IMG, W_IMG = 540, 960
H_PAD, W_PAD = 576, 960          # reflect-padded to a multiple of 64 (datasets/transforms.py:147-161)
MAX_DISP = 192
BATCH = 8
METRIC = "stereo_frames_per_sec_960x540_D192"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the torch-CUDA eager run of the reference model")
    ap.add_argument("--no-full-codd", action="store_true", help="skip the bounded full-CODD (configs[2]) leg")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--dump-kernels", default=None, help="write the per-kernel table of the instrumented pass (JSON)")
    return ap.parse_args()


def workload_config(batch, n_gpus):
    return {
        "workload": f"HITNetMF stereo-only forward, batch={batch} per GPU, 960x540 pairs padded to "
                    f"{H_PAD}x{W_PAD}, D={MAX_DISP}, random-init weights (BASELINE.json configs[1])",
        "per_gpu_batch": batch, "global_batch": batch * n_gpus, "max_disp": MAX_DISP,
        "padded_hw": [H_PAD, W_PAD], "parallelism": f"dp{n_gpus} (batch-sharded, weights broadcast once)",
        # (timing rule: inputs vs L2) — part of the workload description, identical in both arms
        "l2": "activations of one step (several GB) exceed the 126 MB L2; no explicit flush",
    }


# ----------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi in the background during the timed region)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5),
                              ("sw_power_cap", 6)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median of the samples taken under load (upper half: idle samples bracket the region)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# reference / CPU arm
# ----------------------------------------------------------------------------------------------
def reference_forward(device="cpu"):
    """-> (fn(left, right) -> pred_disp, kind, what).  The unmodified reference HITNetMF when its mirror travels with
    the snapshot (baseline/_ref), else the oracle port in reference form.  Same weights either way (seed 0)."""
    import contextlib
    from oracle import hitnet_oracle as O
    sd = O.random_hitnet_params(0)
    try:
        from oracle import ref_loader
        if ref_loader.available():
            with contextlib.redirect_stdout(sys.stderr):
                m = ref_loader.build_hitnet(MAX_DISP)
            m.load_state_dict(sd, strict=True)
            m.to(device).eval()
            return (lambda l, r: m.stereo_matching(l, r)["pred_disp"]), "reference", \
                "unmodified reference HITNetMF.stereo_matching (model/stereo/hitnet/hitnet.py:75-100, mirror in baseline/_ref)"
    except Exception as exc:  # noqa: BLE001
        print(f"bench: reference mirror unusable ({type(exc).__name__}: {exc}); timing the oracle port", file=sys.stderr)
    sd = {k: v.to(device) for k, v in sd.items()}
    return (lambda l, r: O.stereo_matching(sd, l, r, MAX_DISP, reference_form=True)["pred_disp"]), "port", \
        "oracle port in reference form (5-D grid_sample cost volume)"


def cpu_forward_time(rows, repeats, fwd=None):
    """Seconds per pair for the CPU arm on a [1,3,rows,960] strip (rows % 64 == 0)."""
    from oracle import hitnet_oracle as O
    fwd = fwd or reference_forward("cpu")[0]
    left, right = O.synth_pair(1, rows, W_PAD, MAX_DISP, seed=1234, kind="S")
    ts = []
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            fwd(left, right)
            ts.append(time.perf_counter() - t0)
    return ts


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import hitnet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    fwd, kind, what = reference_forward("cpu")
    # bounded sample: one full 576x960 pair per step if the run fits ~4 minutes, else a strip of rows
    probe = cpu_forward_time(128, 2, fwd)[1]               # 128 rows: smallest valid strip
    est_full = probe * (H_PAD / 128.0)
    total = args.steps + args.warmup
    rows = H_PAD
    while rows > 128 and est_full * (rows / H_PAD) * total > 240.0:
        rows -= 64
    left, right = O.synth_pair(1, rows, W_PAD, MAX_DISP, seed=1234, kind="S")
    with torch.no_grad():
        for _ in range(args.warmup):
            fwd(left, right)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fwd(left, right)
        dt = time.perf_counter() - t0
    frames = args.steps * rows / H_PAD
    value = frames / dt
    sample = (f"{args.steps} steps x 1 pair of {rows}x{W_PAD} rows ({rows / H_PAD:.3f} frame each; the metric is per "
              f"frame), {what}, torch {torch.__version__} CPU, {torch.get_num_threads()} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.batch, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def gpu_eager_baseline(dev, batch, steps):
    """The unmodified reference model run by torch-CUDA eager (cuDNN sm_100 kernels) on this B200, same batch and
    size: protocol of /root/reference/benchmark_speed.py:44-64 (warm-up, then timed forwards), CUDA events."""
    from oracle import hitnet_oracle as O
    out = {"unit": UNIT, "batch": batch, "steps": steps}
    try:
        fwd, kind, what = reference_forward(dev)
        out["kind"], out["what"] = kind, what + ", torch-CUDA eager on the same GPU"
        left, right = O.synth_pair(batch, H_PAD, W_PAD, MAX_DISP, seed=1234, kind="S")
        left, right = left.to(dev), right.to(dev)
        for flag, name in ((False, "tf32_off"), (True, "tf32_on")):
            torch.backends.cudnn.allow_tf32 = flag
            torch.backends.cuda.matmul.allow_tf32 = flag
            with torch.no_grad():
                for _ in range(2):
                    fwd(left, right)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    fwd(left, right)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"value": round(batch / (ms * 1e-3), 2), "ms_per_step": round(ms, 3)}
        out["value"] = max(out["tf32_off"]["value"], out["tf32_on"]["value"])
        out["peak_mem_gb"] = round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 2)
    except Exception as exc:  # noqa: BLE001   (a baseline leg must not take the headline numbers down)
        out["error"] = f"{type(exc).__name__}: {exc}"
    finally:
        torch.backends.cudnn.allow_tf32 = True
        torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
def pin_to_gpu_numa(index):
    """Bind this rank's threads (and so its pinned host buffers, by first touch) to the CPUs NVML reports as local to
    its GPU; with 8 ranks on one box the fp32 e2e leg is bounded by host memory traffic, and a rank whose staging
    buffers sit on the other socket pays for it (e2e scaling 0.888 at 8 GPUs)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{min(cpus)}-{max(cpus)} ({len(cpus)} cpus, NVML affinity of GPU {index})"
    except Exception as exc:  # noqa: BLE001
        return f"unchanged ({type(exc).__name__})"
    return "unchanged"


def full_codd_leg(dev, reps=3):
    """BASELINE.json configs[2]: full CODD (stereo + motion + fusion) 2-frame forward, batch 4, 576x960, D=192, RAFT3D
    iters=16, through model(...) (reference: model/codd.py:80-126).  Bounded: 2 warm-up + `reps` timed sequences."""
    import codd_b200
    from codd_b200 import ops
    from codd_b200.synth import synth_pair
    out = {"workload": "full CODD 2-frame forward, batch 4, 960x540 (576x960), D=192, iters=16 (BASELINE.json configs[2])"}
    try:
        B, iters = 4, 16
        torch.manual_seed(0)
        model = codd_b200.build_estimator(codd_b200.codd_full_config(MAX_DISP, iters)).to(dev)
        model.eval()
        left, right = synth_pair(B, H_PAD, W_PAD, MAX_DISP, seed=1234, kind="S")
        img = torch.stack([left, torch.roll(left, shifts=(1, 2), dims=(2, 3))], 1).to(dev)
        r_img = torch.stack([right, torch.roll(right, shifts=(1, 2), dims=(2, 3))], 1).to(dev)
        metas = [[dict(min_disp=1, max_disp=MAX_DISP, ori_shape=(H_IMG, W_IMG), img_shape=(H_IMG, W_IMG),
                       intrinsics=[1050.0, 1050.0, W_IMG / 2.0, H_IMG / 2.0])]]

        def run():
            return model(return_loss=False, rescale=True, evaluate=False, img=[img], img_metas=metas, r_img=[r_img])[0]

        with torch.no_grad():
            for _ in range(2):
                res = run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                res = run()
            e1.record()
            torch.cuda.synchronize()
            sec = e0.elapsed_time(e1) * 1e-3 / reps
            n0 = ops.LAUNCHES[0]
            with ops.profile() as prof:
                run()
            kernels = prof.summary()
        total = sum(k["ms"] for k in kernels)
        out.update({"value": round(2 * B / sec, 2), "unit": "frames/s", "seconds_per_sequence": round(sec, 4),
                    "finite": bool(torch.isfinite(res).all()), "launches_per_sequence": ops.LAUNCHES[0] - n0,
                    "kernel_ms_per_sequence": round(total, 2),
                    "top_kernels": [{"kernel": k["kernel"], "ms": round(k["ms"], 3), "launches": k["launches"],
                                     "gbs": round(k["bytes"] / (k["ms"] * 1e-3) / 1e9, 1) if k["bytes"] else None}
                                    for k in kernels[:8]]})
        del model
    except Exception as exc:  # noqa: BLE001
        out["error"] = f"{type(exc).__name__}: {exc}"
    torch.cuda.empty_cache()
    return out


def run_b200(args):
    import torch.distributed as dist
    import codd_b200
    from codd_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: codd_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = pin_to_gpu_numa(local)       # before any pinned host allocation (first touch decides the NUMA node)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # model: random init on rank 0, one flat NCCL broadcast of the weights (the reference's DDP wrap)
    torch.manual_seed(0)
    model = codd_b200.build_estimator(codd_b200.codd_stereo_config(MAX_DISP)).to(dev)
    model.eval()
    from codd_b200.sharding import broadcast_parameters, reduce_max_ms
    broadcast_parameters(model, src=0)

    B = args.batch
    # Set U of SURVEY §8d would feed right == left (degenerate ties); use independent textured pairs
    from codd_b200.synth import synth_pair
    left_h, right_h = synth_pair(B, H_PAD, W_PAD, MAX_DISP, seed=1234 + rank, kind="S")
    left_h, right_h = left_h.pin_memory(), right_h.pin_memory()
    left, right = left_h.to(dev), right_h.to(dev)
    stereo = model.stereo

    def step():
        return stereo.stereo_matching(left, right)["pred_disp"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            out = step()
        torch.cuda.synchronize()
        n0 = ops.LAUNCHES[0]
        step()
        launches_per_step = ops.LAUNCHES[0] - n0

        graph = None
        if not args.no_graph:
            # the ~130 launches of a step are captured once and replayed (static shapes)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = step()

        def run_step():
            if graph is not None:
                graph.replay()
            else:
                step()

        for _ in range(3):
            run_step()

        # ---------------- timed region: device-resident inputs
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
            time.sleep(0.3)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            run_step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None
        ms = reduce_max_ms(ms, dev)

        # ---------------- e2e: public API, pinned host buffers, H2D + D2H inside the region
        img_h = torch.stack([left_h], 1).pin_memory()       # [B, MF=1, 3, H, W]
        rimg_h = torch.stack([right_h], 1).pin_memory()
        metas = [[dict(min_disp=1, max_disp=MAX_DISP, ori_shape=(H_IMG, W_IMG), img_shape=(H_IMG, W_IMG))]]
        res_h = torch.empty((B, 1, H_IMG, W_IMG), dtype=torch.float32).pin_memory()

        # two serving streams, used alternately: the H2D copy of step i+1 and the D2H read of step i-1 overlap the
        # kernels of step i (every step still does its own H2D -> model(...) -> D2H, in order, on its stream)
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        res_hs = [res_h, torch.empty_like(res_h).pin_memory()]

        def e2e_step(i):
            with torch.cuda.stream(streams[i % 2]):
                img = img_h.to(dev, non_blocking=True)
                rimg = rimg_h.to(dev, non_blocking=True)
                res = model(return_loss=False, rescale=True, evaluate=False, img=[img], img_metas=metas, r_img=[rimg])
                res_hs[i % 2].copy_(res[0], non_blocking=True)

        def e2e_join():
            for st in streams:
                torch.cuda.current_stream().wait_stream(st)

        e2e_steps = max(4, min(args.steps, 50))     # the first H2D and the last D2H are not overlapped: amortised over the steps
        for st in streams:
            st.wait_stream(torch.cuda.current_stream())
        for i in range(2):
            e2e_step(i)
        e2e_join()
        barrier()
        e0.record()
        for st in streams:
            st.wait_stream(torch.cuda.current_stream())
        for i in range(e2e_steps):
            e2e_step(i)
        e2e_join()
        e1.record()
        barrier()
        e2e_ms = e0.elapsed_time(e1)
        e2e_ms = reduce_max_ms(e2e_ms, dev)
        h2d = img_h.numel() * 4 + rimg_h.numel() * 4

        # ---------------- e2e with the input-staging step on the GPU: uint8 frames cross PCIe, the
        # Normalize + reflect Pad(64) + HWC->CHW of the reference's CPU pipeline runs as one kernel per view
        g8 = torch.Generator().manual_seed(99 + rank)
        l8_h = torch.randint(0, 256, (B, H_IMG, W_IMG, 3), dtype=torch.uint8, generator=g8).pin_memory()
        r8_h = torch.randint(0, 256, (B, H_IMG, W_IMG, 3), dtype=torch.uint8, generator=g8).pin_memory()

        def e2e_u8_step(i):
            with torch.cuda.stream(streams[i % 2]):
                l8 = l8_h.to(dev, non_blocking=True)
                r8 = r8_h.to(dev, non_blocking=True)
                img = ops.stage_images_u8(l8).unsqueeze(1)
                rimg = ops.stage_images_u8(r8).unsqueeze(1)
                res = model(return_loss=False, rescale=True, evaluate=False, img=[img], img_metas=metas, r_img=[rimg])
                res_hs[i % 2].copy_(res[0], non_blocking=True)

        for st in streams:
            st.wait_stream(torch.cuda.current_stream())
        for i in range(2):
            e2e_u8_step(i)
        e2e_join()
        barrier()
        e0.record()
        for st in streams:
            st.wait_stream(torch.cuda.current_stream())
        for i in range(e2e_steps):
            e2e_u8_step(i)
        e2e_join()
        e1.record()
        barrier()
        e2e_u8_ms = reduce_max_ms(e0.elapsed_time(e1), dev)
        h2d_u8 = l8_h.numel() + r8_h.numel()
        d2h = res_h.numel() * 4

        # ---------------- the same through the sequence driver: staging + stereo forward + crop replayed
        # from one CUDA graph per serving slot, uint8 frames in pinned host memory, results in pinned host memory
        from codd_b200.runner import StereoSequenceRunner
        e2e_graph_err = None
        try:
            runner = StereoSequenceRunner(model, device=dev, use_graph=not args.no_graph, n_streams=2)
            for _ in runner.infer_batches([(l8_h, r8_h)] * 2):
                pass
        except Exception as exc:      # an auxiliary leg must not take the headline numbers down with it
            e2e_graph_err = f"{type(exc).__name__}: {exc}"
        barrier()
        e0.record()
        if e2e_graph_err is None:
            try:
                n_out = sum(o.shape[0] for o in runner.infer_batches([(l8_h, r8_h)] * e2e_steps))
                assert n_out == B * e2e_steps
            except Exception as exc:
                e2e_graph_err = f"{type(exc).__name__}: {exc}"
        e1.record()
        barrier()
        e2e_graph_ms = reduce_max_ms(e0.elapsed_time(e1), dev)

        # ---------------- instrumented pass: per-launch events -> dominant kernel + K1/K4 numbers
        with ops.profile() as prof:
            step()
        kernels = prof.summary()
        # the HBM-bound, materialising variant of K1 (training API: the volumes are kept) on the tiles of this step;
        # not part of the eval step, reported beside the fused variant in roofline_named_kernels
        fl, fr = stereo.backbone.forward_pair(left, right)
        tiles = stereo.tile_init.tile_features(fl, fr)
        dks = [MAX_DISP // (16 >> k) for k in range(5)]
        for _ in range(2):
            ops.cost_volume_pyramid(tiles, dks, want_cv=True)
        with ops.profile() as prof2:
            for _ in range(5):
                ops.cost_volume_pyramid(tiles, dks, want_cv=True)
        k1m = prof2.summary()
        del fl, fr, tiles

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    def _flatten(obj, prefix=""):
        out = {}
        if isinstance(obj, dict):
            for k, v in obj.items():
                out.update(_flatten(v, f"{prefix}{k}_" if isinstance(v, dict) else f"{prefix}{k}"))
        elif isinstance(obj, (int, float)) and not isinstance(obj, bool):
            out[prefix] = obj
        return out

    peaks = _flatten(peaks)      # tolerate nesting ({"hbm": {"gbs": ..}} -> "hbm_gbs") and non-dict files
    # the kernels are timed inside a long step: prefer a sustained figure when the driver's file has one
    peak_key = next((k for k in ("hbm_gbs_sustained", "hbm_sustained_gbs", "hbm_gbs", "hbm_gbs_burst")
                     if isinstance(peaks.get(k), (int, float))), None)
    if peak_key is None:
        peak_key = next((k for k, v in peaks.items()           # any HBM figure that can be a bandwidth in GB/s
                         if "hbm" in k.lower() and isinstance(v, (int, float)) and 1000.0 <= v <= 20000.0), None)
    peak_gbs = float(peaks[peak_key]) if peak_key else 6650.0
    peak_src = f"MEASURED_PEAKS.json {peak_key} (measured)" if peak_key else "fallback 6650 GB/s"

    total_ms = sum(k["ms"] for k in kernels)
    # dram__bytes_read + dram__bytes_write per launch from the committed ncu capture of one step
    # (tools/traffic_table.py -> profiles/traffic_r01.json); None for a kernel the capture does not hold
    traffic, traffic_src = {}, None
    for name in ("traffic_r02.json", "traffic_r01.json"):
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", name)))["kernels"]
            traffic_src = "profiles/" + name
            break
        except Exception:  # noqa: BLE001
            continue

    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    # FP32 roof of the fused arg-min cost volume: 2 * 16 * D lane-adds per tile (one FADD for the difference, one for
    # the |.| accumulate; no FMA, the channel sum is sequential for bit-exactness), 128 fp32 lanes per SM and clock
    FP32_LANES = 148 * 128

    def k1_lane_adds(batch):
        tot = 0
        for lv in range(5):
            h, w, d = H_PAD // 4 >> lv, W_PAD // 4 >> lv, MAX_DISP >> lv
            # pairs (j, d) with 4j - d >= 0 do arithmetic; the zero-filled shifts are one |L|_1 per column
            pairs = sum(min(d, 4 * j + 1) for j in range(w)) * h
            tot += 2 * 16 * pairs
        return tot * batch

    def roof(k, brief=False):
        gbs = k["bytes"] / (k["ms"] * 1e-3) / 1e9
        tr = traffic.get(k["kernel"])
        r = {"kernel": k["kernel"], "bound": "hbm", "achieved": round(gbs, 1), "peak": peak_gbs, "unit": "GB/s",
             "frac": round(gbs / peak_gbs, 4),
             "traffic": None if tr is None else round(tr["traffic_per_launch"]),
             "algorithmic_bytes_per_launch": round(k["bytes"] / k["launches"]),
             "launches_per_step": k["launches"],
             "ms_per_step": round(k["ms"], 4), "share_of_step": round(k["ms"] / total_ms, 4)}
        if k["kernel"].startswith("cost_volume") and k["kernel"].endswith("_argmin") and "build" not in k["kernel"]:
            # fused arg-min variant: arithmetic intensity 6144 lane-adds per 328 B, so the FP32 pipe is the roof
            lane_ops = k1_lane_adds(B) * k["launches"]
            peak = FP32_LANES * sm_mhz * 1e6 / 1e12
            ach = lane_ops / (k["ms"] * 1e-3) / 1e12
            r.update({"bound": "fp32", "achieved": round(ach, 2), "peak": round(peak, 2), "unit": "Tlane-op/s",
                      "frac": round(ach / peak, 4), "hbm_gbs": round(gbs, 1), "hbm_frac": round(gbs / peak_gbs, 4),
                      "lane_adds_per_launch": lane_ops // k["launches"],
                      "peak_note": f"148 SMs x 128 fp32 lanes x {sm_mhz:.0f} MHz (median SM clock of this run)"})
        if not brief:
            r.update({"algorithmic_bytes_per_step": k["bytes"], "peak_source": peak_src, "traffic_source": traffic_src,
                      "how": "per-launch CUDA events on the launching stream, instrumented eager pass after the timed region"})
        return r

    if args.dump_kernels:
        json.dump([roof(k) for k in kernels], open(args.dump_kernels, "w"), indent=1)
    dominant = roof(kernels[0])
    named = [roof(k, brief=True) for k in kernels if k["kernel"].startswith(("cost_volume", "tile_warp_cost"))]
    for k in k1m:       # 5 launches outside the step: per-launch figures are what matter
        r = roof(k, brief=True)
        r.update({"launches_per_step": 0, "ms_per_launch": round(k["ms"] / k["launches"], 4), "share_of_step": 0.0,
                  "note": "materialising variant (training API), timed on this step's tile features, 5 launches"})
        del r["ms_per_step"]
        named.append(r)

    frames = args.steps * B * world
    value = frames / (ms * 1e-3)
    e2e_f32 = {"value": round(e2e_steps * B * world / (e2e_ms * 1e-3), 3), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
               "what": "fp32 padded frames cross PCIe (the reference's CPU pipeline output), then model(...)"}
    e2e_u8 = {"value": round(e2e_steps * B * world / (e2e_u8_ms * 1e-3), 3), "unit": UNIT,
              "h2d_bytes_per_step": h2d_u8, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
              "api": "ops.stage_images_u8 (SURVEY 8f N1: Normalize + reflect Pad(64) + HWC->CHW of datasets/transforms.py "
                     "on the GPU, bit-identical to the CPU pipeline) -> ConsistentOnlineDynamicDepth.__call__("
                     "return_loss=False, evaluate=False, img=[..], r_img=[..]); uint8 HWC frames in pinned host memory",
              "pipelining": "2 CUDA streams used alternately (copies of one step overlap kernels of the other)"}
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "e2e": e2e_u8,
        "e2e_f32": e2e_f32,
        "e2e_graph": {"error": e2e_graph_err} if e2e_graph_err else {
                      "value": round(e2e_steps * B * world / (e2e_graph_ms * 1e-3), 3), "unit": UNIT,
                      "h2d_bytes_per_step": h2d_u8, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                      "what": "codd_b200.runner.StereoSequenceRunner (SURVEY 8f N2): staging + stereo + crop replayed "
                              "from one CUDA graph per serving slot, 2 slots"},
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step,
        "config": workload_config(B, world),       # the same dict the reference arm prints
        "run": {"cuda_graph": graph is not None, "cpu_affinity": affinity},
        "clocks": clocks,
    }
    if world == 1 and not args.no_gpu_baseline:
        line["gpu_eager_baseline"] = gpu_eager_baseline(dev, B, max(3, min(args.steps, 10)))
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        fwd, kind, what = reference_forward("cpu")
        ts = cpu_forward_time(H_PAD, 3, fwd)                       # 1 warm-up + 2 timed pairs (~10-20 s)
        sec = sum(ts[1:]) / len(ts[1:])
        line["cpu_baseline"] = {
            "value": round(1.0 / sec, 4), "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"2 timed forwards of 1 pair {H_PAD}x{W_PAD} D={MAX_DISP} (1/{B} of a step) after 1 warm-up; {what}",
        }
    if world == 1 and not args.no_full_codd:
        line["full_codd"] = full_codd_leg(dev)
    dominant["note"] = ("dominant kernel of the step by time; in round 1 this was the 16-channel ring conv (0.52 of HBM), whose "
                        "pairs now run fused (conv3x3x2ring_c16) — the 32-channel ring conv is bound by the L1/shared-memory "
                        "data pipe (tensor-core operand reads + epilogue stores), see DESIGN.md")
    line["roofline"] = dominant
    # the whole step against the same roof: algorithmic bytes of every launch of one step / graph-replayed step time
    step_bytes = sum(k["bytes"] for k in kernels)
    step_gbs = step_bytes / (ms / args.steps * 1e-3) / 1e9
    line["roofline_step"] = {"bound": "hbm", "achieved": round(step_gbs, 1), "peak": peak_gbs, "unit": "GB/s",
                             "frac": round(step_gbs / peak_gbs, 4), "algorithmic_bytes_per_step": step_bytes,
                             "what": "sum of the algorithmic bytes of all launches of one step / timed ms_per_step"}
    line["roofline_named_kernels"] = named
    line["top_kernels"] = [{"kernel": k["kernel"], "ms": round(k["ms"], 4), "launches": k["launches"],
                            "share": round(k["ms"] / total_ms, 4)} for k in kernels[:8]]
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
