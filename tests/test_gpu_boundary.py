"""Boundary hardening: the C ABI under concurrent first calls from several host threads / streams,
on a second device of the same process, and the ring convolution at the edge of its documented fp16 operand range."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = textwrap.dedent('''
    import sys, threading, torch
    sys.path.insert(0, %(root)r)
    from codd_b200 import ops
    from codd_b200.lib import ACT_LEAKY
    devices = %(devices)s
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 32, 40, 200, generator=g)
    wt = torch.randn(32, 32, 3, 3, generator=g) / 17.0
    b = torch.randn(32, generator=g)
    fl = torch.randn(2, 16, 32, 96, generator=g); fr = torch.randn(2, 16, 32, 96, generator=g)
    cur = torch.randn(2, 16, 8, 24, generator=g); cur[:, 0] = torch.rand(2, 8, 24, generator=g) * 20
    prev = torch.randn(2, 16, 4, 12, generator=g); prev[:, 0] = torch.rand(2, 4, 12, generator=g) * 10
    dw = torch.randn(16, 64, generator=g) / 8; db = torch.randn(16, generator=g)
    results, errors = {}, []

    def work(tid, dev):
        # the FIRST call of every kernel in this process happens here, concurrently in all threads: the per-device
        # attribute set-up (CoddDeviceOnce) must neither race nor skip a device
        try:
            torch.cuda.set_device(dev)
            st = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(st), torch.no_grad():
                d = torch.device("cuda", dev)
                for it in range(3):
                    y = ops.conv3x3_tc_ring(ops.to_nhwc(x.to(d)), ops.pack_conv_weight_ring(wt.to(d)), b.to(d), 32, ACT_LEAKY)
                    aug = ops.tile_warp_cost(ops.to_nhwc(fl.to(d)), ops.to_nhwc(fr.to(d)), ops.to_nhwc(cur.to(d)),
                                             ops.to_nhwc(prev.to(d)), dw.to(d).contiguous(), db.to(d))
                st.synchronize()
                results[tid] = (ops.to_nchw(y).cpu(), ops.to_nchw(aug).cpu())
        except Exception as exc:
            errors.append(f"thread {tid} dev {dev}: {type(exc).__name__}: {exc}")

    threads = [threading.Thread(target=work, args=(i, devices[i %% len(devices)])) for i in range(%(nthreads)d)]
    [t.start() for t in threads]; [t.join() for t in threads]
    assert not errors, errors
    ref = results[0]
    for tid, r in results.items():
        assert torch.equal(r[0], ref[0]) and torch.equal(r[1], ref[1]), f"thread {tid} differs"
    import torch.nn.functional as F
    want = F.leaky_relu(F.conv2d(x, wt, b, padding=1), 0.2)
    torch.testing.assert_close(ref[0], want, rtol=2e-5, atol=2e-5)
    print("OK", len(results), "threads on devices", devices)
''')


def _run_worker(devices, nthreads):
    code = _WORKER % dict(root=ROOT, devices=repr(devices), nthreads=nthreads)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout


def test_concurrent_first_calls_from_four_threads():
    """Fresh process, four host threads with their own streams make the first ring-conv / K4 calls at the same time."""
    _run_worker([0], 4)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process():
    """Round 1 kept one `static bool configured` per kernel: a second device never got its shared-memory attribute."""
    _run_worker([0, 1], 4)


def test_ring_conv_at_the_fp16_operand_range():
    """The ring kernel splits operands into fp16 halves: |x|, |w| up to 65504 are exact-range inputs (result fp32-class
    relative to the output scale), larger magnitudes are CLAMPED to 65504 as include/codd_b200.h documents."""
    from codd_b200 import ops
    from codd_b200.lib import ACT_NONE
    g = torch.Generator().manual_seed(9)
    x = (torch.rand(1, 16, 12, 140, generator=g) * 2 - 1) * 65504.0
    x[0, :, 3, 10:20] = 65504.0
    x[0, :, 4, 10:20] = -65504.0
    wt = torch.randn(16, 16, 3, 3, generator=g) / 12.0
    b = torch.randn(16, generator=g)
    ref = F.conv2d(x.double(), wt.double(), b.double(), padding=1).float()
    out = ops.conv3x3_tc_ring(ops.to_nhwc(x.cuda()), ops.pack_conv_weight_ring(wt.cuda()), b.cuda(), 16, ACT_NONE)
    got = ops.to_nchw(out).cpu()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    print(f"ring conv at |x| <= 65504: max abs err {err:.3e} on outputs of scale {scale:.3e} (rel {err / scale:.2e})")
    assert torch.isfinite(got).all() and err <= 2e-5 * scale
    # beyond the range: clamped, finite, equal to the convolution of the clamped input
    x2 = x.clone()
    x2[0, :, 6, 30:40] = 1.0e6
    x2[0, :, 7, 30:40] = -3.0e5
    ref2 = F.conv2d(x2.clamp(-65504.0, 65504.0).double(), wt.double(), b.double(), padding=1).float()
    got2 = ops.to_nchw(ops.conv3x3_tc_ring(ops.to_nhwc(x2.cuda()), ops.pack_conv_weight_ring(wt.cuda()), b.cuda(), 16,
                                           ACT_NONE)).cpu()
    assert torch.isfinite(got2).all() and (got2 - ref2).abs().max().item() <= 2e-5 * ref2.abs().max().item()
