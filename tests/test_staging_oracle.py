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
