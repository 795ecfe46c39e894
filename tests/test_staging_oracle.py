"""N1 oracle (oracle/staging_oracle.py): its pieces against independent statements (numpy only, CPU)."""
import numpy as np

from oracle import staging_oracle as SO


def test_reflect_pad_and_normalisation():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(1, 70, 130, 3), dtype=np.uint8)
    out = SO.stage_images_u8(img)
    assert out.shape == (1, 3, 128, 192) and out.dtype == np.float32
    # channel 0 of the output is R = input channel 2 (BGR frames, to_rgb=True)
    r = (img[0, :, :, 2].astype(np.float32) - np.float32(123.675)) * (np.float32(1) / np.float32(58.395))
    assert np.array_equal(out[0, 0, :70, :130], r)
    # reflect without repeating the edge: row 70 mirrors row 68, column 130 mirrors column 128
    assert np.array_equal(out[0, :, 70, :130], out[0, :, 68, :130])
    assert np.array_equal(out[0, :, :70, 130], out[0, :, :70, 128])
    assert np.array_equal(out[0, :, 127, 191], out[0, :, 2 * 69 - 127, 2 * 129 - 191])
    # already aligned sizes are not padded
    assert SO.stage_images_u8(img[:, :64, :128]).shape == (1, 3, 64, 128)


def test_against_the_opencv_calls_mmcv_makes():
    """mmcv itself is absent, but OpenCV is here: `mmcv.imnormalize` is cv2.cvtColor(BGR2RGB) + cv2.subtract(img, mean)
    + cv2.multiply(img, 1/std) on a float32 image with float64 (1,3) mean / reciprocal arrays, and `mmcv.impad(...,
    padding_mode='reflect')` is cv2.copyMakeBorder(BORDER_REFLECT_101) (mmcv 1.7.0 image/photometric.py, geometric.py).
    The restatement matches those calls to one float32 ulp (cv2.multiply works with the float64 reciprocal, the oracle
    and the CUDA kernel with its float32 rounding); the padding geometry is identical."""
    import pytest
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(70, 130, 3), dtype=np.uint8)
    mean = np.float64(np.array([123.675, 116.28, 103.53]).reshape(1, -1))
    stdinv = 1 / np.float64(np.array([58.395, 57.12, 57.375]).reshape(1, -1))
    x = img.astype(np.float32).copy()
    cv2.cvtColor(x, cv2.COLOR_BGR2RGB, x)
    cv2.subtract(x, mean, x)
    cv2.multiply(x, stdinv, x)
    x = cv2.copyMakeBorder(x, 0, 128 - 70, 0, 192 - 130, cv2.BORDER_REFLECT_101)
    ref = x.transpose(2, 0, 1)
    out = SO.stage_images_u8(img[None])[0]
    assert out.shape == ref.shape
    ulp = np.abs(out - ref) / np.maximum(np.spacing(np.abs(ref).astype(np.float32)), np.float32(1e-12))
    assert ulp.max() <= 1.0, ulp.max()
    # the padded region is a pure copy of interior pixels in both
    assert np.array_equal(out[:, 70:, :130] == out[:, 68:10:-1, :130][:, :58], np.ones_like(out[:, 70:, :130], bool))
