"""N4: on-disk formats (PFM, KITTI 16-bit PNG disparity / flow, reference checkpoints, .disp.pred.npz).  CPU only."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from codd_b200 import io as cio
from oracle import ref_loader


def write_pfm(path, arr, little=True, scale=1.0):
    arr = np.asarray(arr, np.float32)
    with open(path, "wb") as f:
        f.write(b"PF\n" if arr.ndim == 3 else b"Pf\n")
        f.write(f"{arr.shape[1]} {arr.shape[0]}\n".encode())
        f.write(f"{-scale if little else scale}\n".encode())
        f.write(np.flipud(arr).astype("<f4" if little else ">f4").tobytes())


def _ref_data_io():
    """The reference's datasets/data_io.py, imported as a file (its package __init__ pulls in mmcv datasets)."""
    if not ref_loader.available():
        return None
    ref_loader.load()   # puts the mmcv shim on sys.path
    spec = importlib.util.spec_from_file_location("_ref_data_io", os.path.join(ref_loader.reference_root(), "datasets", "data_io.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("little", [True, False])
@pytest.mark.parametrize("shape", [(7, 11), (5, 9, 3)])
def test_read_pfm(tmp_path, little, shape):
    arr = np.random.default_rng(0).normal(0, 50, shape).astype(np.float32)
    p = str(tmp_path / "x.pfm")
    write_pfm(p, arr, little=little, scale=2.5)
    data, scale = cio.read_pfm(p)
    assert scale == 2.5 and data.dtype.kind == "f" and np.array_equal(np.asarray(data, np.float32), arr)
    ref = _ref_data_io()
    if ref is not None:                               # pinned against the reference reader
        rdata, rscale = ref.read_pfm(p)
        assert rscale == scale and np.array_equal(rdata, data)


def test_read_pfm_rejects_other_files(tmp_path):
    p = tmp_path / "bad.pfm"
    p.write_bytes(b"P6\n3 3\n255\n" + bytes(27))
    with pytest.raises(ValueError):
        cio.read_pfm(str(p))


def test_read_kitti_disp_and_flow():
    cv2 = pytest.importorskip("cv2")
    g = np.random.default_rng(1)
    disp16 = g.integers(0, 65535, (13, 17), dtype=np.uint16)
    disp16[g.random((13, 17)) < 0.3] = 0
    ok, buf = cv2.imencode(".png", disp16)
    assert ok
    d = cio.read_kitti_disp(buf.tobytes())
    assert d.shape == (13, 17) and np.array_equal(d, disp16 / 256.0)           # data_io.py:226-228
    # flow PNG: channels (R, G, B) = (u, v, valid), stored by cv2 as BGR
    u = g.integers(0, 65535, (13, 17), dtype=np.uint16)
    v = g.integers(0, 65535, (13, 17), dtype=np.uint16)
    val = (g.random((13, 17)) > 0.4).astype(np.uint16)
    ok, buf = cv2.imencode(".png", np.stack([val, v, u], -1))
    flow, valid = cio.read_kitti_flow(buf.tobytes())
    assert flow.dtype == np.float32 and flow.shape == (13, 17, 2)
    assert np.array_equal(flow[..., 0], (u.astype(np.float32) - 2 ** 15) / 64.0)   # data_io.py:231-236
    assert np.array_equal(flow[..., 1], (v.astype(np.float32) - 2 ** 15) / 64.0)
    assert np.array_equal(valid, val.astype(np.float32))


def test_load_checkpoint_reference_layout(tmp_path):
    """A checkpoint written the way mmcv's runner does ({"meta", "state_dict"} with DDP "module." prefixes and the
    reference's parameter names) loads key for key into the drop-in HITNetMF."""
    import codd_b200
    torch.manual_seed(3)
    src = codd_b200.MODELS.build(codd_b200.hitnet_config(64))
    sd = {"module." + k: v.clone() + 0.25 for k, v in src.state_dict().items()}
    p = str(tmp_path / "ckpt.pth")
    torch.save({"meta": {"epoch": 1}, "state_dict": sd}, p)
    dst = codd_b200.MODELS.build(codd_b200.hitnet_config(64))
    missing, unexpected = cio.load_checkpoint(dst, p)
    assert missing == [] and unexpected == []
    for (k, a), (_, b) in zip(dst.state_dict().items(), src.state_dict().items()):
        assert torch.equal(a, b + 0.25), k
    # bare state_dict files work too; unknown keys are reported, not fatal (strict=False as in the reference)
    torch.save({**src.state_dict(), "extra.weight": torch.zeros(1)}, p)
    missing, unexpected = cio.load_checkpoint(dst, p)
    assert missing == [] and unexpected == ["extra.weight"]


def test_write_disp_npz(tmp_path):
    disp = np.random.default_rng(2).random((1, 1, 6, 8)).astype(np.float32)
    target = cio.write_disp_npz(str(tmp_path / "out" / "frame_0001.png"), disp)
    assert target.endswith("frame_0001.disp.pred.npz")
    assert np.array_equal(np.load(target)["disp"], disp)
