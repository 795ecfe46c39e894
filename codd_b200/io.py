"""On-disk formats at the edges of the hot path, without the mm* stack.

Host code by nature (the reference's is numpy + cv2 through mmcv):
  read_pfm          datasets/data_io.py:239-285   (SceneFlow / FlyingThings disparity and flow)
  read_kitti_disp   datasets/data_io.py:226-228   (16-bit PNG, disparity * 256)
  read_kitti_flow   datasets/data_io.py:231-236   (16-bit 3-channel PNG, (v - 2^15) / 64 + validity)
  load_checkpoint   inference.py:123 (mmcv.runner.load_checkpoint(model, path, map_location="cpu")): a .pth holding either a
                    bare state_dict or {"state_dict": ..., "meta": ...}, keys optionally prefixed with "module."
  write_disp_npz    the `.disp.pred.npz` result files of model/codd.py:596-599
The modules of this package keep the reference's parameter names, so reference checkpoints load key for key.
"""
import re

import numpy as np


def read_pfm(path):
    """-> (data float32 [H,W] or [H,W,3], bottom-up rows flipped to top-down, scale)."""
    with open(path, "rb") as f:
        header = f.readline().rstrip().decode("ascii")
        if header == "PF":
            color = True
        elif header == "Pf":
            color = False
        else:
            raise ValueError("Not a PFM file: " + str(path))
        m = re.match(r"^(\d+)\s(\d+)\s$", f.readline().decode("ascii"))
        if not m:
            raise ValueError("Malformed PFM header.")
        width, height = int(m.group(1)), int(m.group(2))
        scale = float(f.readline().decode("ascii").rstrip())
        endian = "<" if scale < 0 else ">"      # negative scale = little-endian samples
        data = np.frombuffer(f.read(), endian + "f")
    data = np.flipud(data.reshape((height, width, 3) if color else (height, width)))
    return data, abs(scale)


def _decode_png_unchanged(img_bytes):
    """PNG bytes -> array with the file's own depth and channel order as cv2 returns it (BGR)."""
    try:
        import cv2
        arr = cv2.imdecode(np.frombuffer(img_bytes, np.uint8), cv2.IMREAD_UNCHANGED)
        if arr is None:
            raise ValueError("could not decode image bytes")
        return arr
    except ImportError:      # Pillow decodes 16-bit gray; 16-bit RGB needs cv2
        import io
        from PIL import Image
        arr = np.array(Image.open(io.BytesIO(img_bytes)))
        if arr.ndim == 3:
            if arr.dtype != np.uint16:
                raise ValueError("16-bit colour PNGs need OpenCV")
            arr = arr[:, :, ::-1]
        return arr


def read_kitti_disp(img_bytes):
    """-> float64 [H,W] disparity (0 = invalid), as the reference's `/ 256.0` on the uint16 image."""
    return _decode_png_unchanged(img_bytes).squeeze() / 256.0


def read_kitti_flow(img_bytes):
    """-> (flow float32 [H,W,2], valid float32 [H,W])."""
    flow = _decode_png_unchanged(img_bytes)[:, :, ::-1].astype(np.float32)      # BGR -> (u, v, valid)
    flow, valid = flow[:, :, :2], flow[:, :, 2]
    return (flow - 2 ** 15) / 64.0, valid


def load_checkpoint(model, path, map_location="cpu", strict=False):
    """Load a reference `.pth` into a codd_b200 (or any torch) module.  Returns (missing_keys, unexpected_keys)."""
    import torch
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    if not isinstance(ckpt, dict):
        raise RuntimeError(f"No state_dict found in checkpoint file {path}")
    state = ckpt.get("state_dict", ckpt)
    state = {(k[7:] if k.startswith("module.") else k): v for k, v in state.items()}
    res = model.load_state_dict(state, strict=strict)
    return list(res.missing_keys), list(res.unexpected_keys)


def write_disp_npz(out_file, disp):
    """The reference's show_result (codd.py:596-599): `<out_file without extension>.disp.pred.npz` with key "disp"."""
    import os
    target = out_file.replace(os.path.splitext(out_file)[1], ".disp.pred.npz")
    os.makedirs(os.path.dirname(target) or ".", exist_ok=True)
    with open(target, "wb") as f:
        np.savez_compressed(f, disp=np.asarray(disp))
    return target
