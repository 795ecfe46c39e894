"""Static checks on the compiled sm_100a objects (cuobjdump, no GPU): the properties that cost the most time to get
right and that a harmless-looking edit can silently undo.  Skipped when the build directory or cuobjdump is missing."""
import os
import re
import shutil
import subprocess

import pytest

BUILD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "codd_b200", "csrc", "build")
needs_build = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(os.path.join(BUILD, "conv_tc_ring.o")),
                                 reason="needs cuobjdump and the built objects (python -c 'import __graft_entry__ as g; g.build()')")


def _sass(obj):
    return subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True, check=True).stdout


def _functions(sass):
    out, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            out[name] = []
        elif name:
            out[name].append(line)
    return out


@needs_build
def test_ring_conv_uses_tcgen05_tma_and_uniform_issue():
    fns = {k: v for k, v in _functions(_sass("conv_tc_ring.o")).items() if "conv3x3_tc_ring_kernel" in k}
    assert len(fns) == 4                                          # <16,16>, <32,16>, <32,32>, <32,32> dilation 3
    for name, lines in fns.items():
        text = "\n".join(lines)
        assert "UTCHMMA" in text, name                            # tcgen05.mma kind::f16
        assert "UTMALDG" in text, name                            # TMA tensor loads
        assert "UTCBAR" in text or "UTCCOMMIT" in text or "ARRIVE" in text, name
        # issuers elected with elect.sync: operands stay in uniform registers.  With `if (lane == 0)` ptxas wrapped every
        # MMA in an elect / R2UR / issue / BRA.U.ANY waterfall loop (79 of them in this kernel, 10 instructions per MMA)
        assert text.count("BRA.U.ANY") <= 4, (name, text.count("BRA.U.ANY"))
        assert text.count("UTCHMMA") >= 40, name                  # the interior paths are fully unrolled


@needs_build
def test_no_stack_frame_in_the_call_free_kernels():
    """Kernels without out-of-line calls must not need a stack frame: accumulators indexed by a run-time loop bound, or
    weights CSE'd across unrolled columns, once put their register arrays in local memory ("What made the
    direct convolutions slow").  (Kernels that call the out-of-line transcendental activations legitimately have one.)"""
    res = ""
    for o in ("conv_tc_ring.o", "gemm_tc.o", "cost_volume.o", "tile_features.o", "metrics.o"):
        res += subprocess.run(["cuobjdump", "-res-usage", os.path.join(BUILD, o)], capture_output=True, text=True,
                              check=True).stdout
    entries = re.findall(r"Function (\S+):\s*\n\s*(.*)", res)
    assert len(entries) >= 15
    for name, usage in entries:
        stack = int(re.search(r"STACK:(\d+)", usage).group(1))
        local = int(re.search(r"LOCAL:(\d+)", usage).group(1))
        assert local == 0 and stack == 0, (name, usage)


@needs_build
def test_round2_kernels_use_the_blackwell_paths():
    """K4 (tile_warp.cu): TMA tensor loads for the window, packed fp32x2 arithmetic in the channel loop, the `decrease`
    tail on tensor-core MMAs.  Strided convolutions / tile features (conv_tc_s2.cu): tcgen05.mma + TMEM loads + 5-D TMA boxes,
    no stack frame."""
    k4 = {k: "\n".join(v) for k, v in _functions(_sass("tile_warp.o")).items() if "tile_warp_cost2_kernel" in k}
    assert len(k4) == 6                                           # NSETS in {1,2} x C in {16,24,32}
    for name, text in k4.items():
        assert "UTMALDG" in text, name                            # cp.async.bulk.tensor window boxes
        assert text.count("FFMA2") >= 24 and "FMUL2" in text and "FADD2" in text, name
        assert text.count("HMMA.1688.F32.TF32") == 12, name       # 2 k-steps x 2 n-tiles x 3 (3xTF32)
        assert "LDS.128" in text, name                            # the window reads stay 128-bit shared loads
    s2 = {k: "\n".join(v) for k, v in _functions(_sass("conv_tc_s2.o")).items() if "conv4x4s2_tc_kernel" in k}
    assert len(s2) == 8
    for name, text in s2.items():
        assert "UTCHMMA" in text or "UTCMMA" in text or re.search(r"UTC\w*MMA", text), name
        assert "UTMALDG.5D" in text or "UTMALDG" in text, name
        assert "LDTM" in text, name
    res = subprocess.run(["cuobjdump", "-res-usage", os.path.join(BUILD, "conv_tc_s2.o")], capture_output=True, text=True,
                         check=True).stdout
    for name, usage in re.findall(r"Function (\S+):\s*\n\s*(.*)", res):
        assert int(re.search(r"STACK:(\d+)", usage).group(1)) == 0, (name, usage)


@needs_build
def test_late_round2_kernels():
    """The fused two-convolution ring kernel (conv_tc_ring2.cu): tcgen05.mma kind::f16 for both convolutions, TMA input
    rows, TMEM loads in both epilogues, 256-bit global accesses, packed fp32x2 epilogue arithmetic, the programmatic
    dependent launch pair (PREEXIT = griddepcontrol.launch_dependents, ACQBULK = griddepcontrol.wait), no stack frame.
    The four tcgen05 kernels carry the dependent-launch pair; the 1x1 convolution moves 256-bit sectors."""
    text = "\n".join(l for v in _functions(_sass("conv_tc_ring2.o")).values() for l in v)
    assert text.count("UTCHMMA") >= 80 and "UTMALDG" in text and text.count("LDTM") >= 4
    assert "STG.E.ENL2.256" in text and "LDG.E.ENL2.256" in text
    assert "FFMA2" in text and "FMUL2" in text and "FADD2" in text
    res = subprocess.run(["cuobjdump", "-res-usage", os.path.join(BUILD, "conv_tc_ring2.o")], capture_output=True, text=True,
                         check=True).stdout
    for name, usage in re.findall(r"Function (\S+):\s*\n\s*(.*)", res):
        assert int(re.search(r"STACK:(\d+)", usage).group(1)) == 0, (name, usage)
        assert int(re.search(r"REG:(\d+)", usage).group(1)) <= 96, (name, usage)       # 672 threads per CTA
    for obj, kern in (("conv_tc_ring.o", "conv3x3_tc_ring_kernel"), ("conv_tc_ring2.o", "conv3x3x2_tc_ring_kernel"),
                      ("conv_tc.o", "conv3x3_tc_kernel"), ("conv_tc_s2.o", "conv4x4s2_tc_kernel")):
        for name, lines in _functions(_sass(obj)).items():
            if kern in name:
                t = "\n".join(lines)
                assert "PREEXIT" in t and "ACQBULK" in t, name
    pw = {k: "\n".join(v) for k, v in _functions(_sass("conv.o")).items() if "pointwise_kernel" in k and "Lb1" in k}
    assert pw
    for name, t in pw.items():
        assert "LDG.E.ENL2.256" in t and "STG.E.ENL2.256" in t, name
