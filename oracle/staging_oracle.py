"""CPU oracle for the input-staging step N1: Normalize -> Pad(size_divisor=64, reflect) -> HWC->CHW of the
reference's test pipeline (datasets/transforms.py:391-421, :147-176; datasets/formating.py:77-85).

TEST INFRASTRUCTURE ONLY.  **Parity unpinned**: the arithmetic lives in mmcv (`mmcv.imnormalize`, `mmcv.impad`,
mmcv-full==1.7.0 per README.md:41), which is absent from /root/reference and from this container; restated from its
published algorithm: float32 image, optional BGR->RGB, subtract mean, multiply by 1/std (reciprocal rounded to fp32
here), cv2.copyMakeBorder(BORDER_REFLECT_101) == numpy 'reflect' on the bottom / right.  tests/test_staging_oracle.py
checks it against the OpenCV calls mmcv 1.7.0 makes (cvtColor / subtract / multiply / copyMakeBorder): identical
geometry, values within one float32 ulp (mmcv multiplies by a float64 reciprocal)."""
import math

import numpy as np


def stage_images_u8(img, mean=(123.675, 116.28, 103.53), std=(58.395, 57.12, 57.375), to_rgb=True, size_divisor=64):
    """img uint8 [N,H,W,3] -> float32 [N,3,Hp,Wp]."""
    x = np.asarray(img).astype(np.float32)
    if to_rgb:
        x = x[..., ::-1]
    mean = np.asarray(mean, np.float32).reshape(1, 1, 1, 3)
    stdinv = (np.float32(1.0) / np.asarray(std, np.float32)).reshape(1, 1, 1, 3)
    x = (x - mean) * stdinv
    n, h, w, _ = x.shape
    hp, wp = math.ceil(h / size_divisor) * size_divisor, math.ceil(w / size_divisor) * size_divisor
    x = np.pad(x, ((0, 0), (0, hp - h), (0, wp - w), (0, 0)), mode="reflect")
    return np.ascontiguousarray(x.transpose(0, 3, 1, 2))
