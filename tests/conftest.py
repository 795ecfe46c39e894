import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are SKIPPED (not failed) on a machine without a CUDA device or without the built library, so a
    plain `pytest tests` goes green on host-only CI.  On a GPU box nothing is skipped here: a missing .so fails loudly."""
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200): run with `pytest -m gpu` on the GPU box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_small():
    return load_golden("hitnet_s_128x128_d32.npz")


@pytest.fixture(scope="session")
def golden_big():
    return load_golden("hitnet_s_128x192_d64.npz")


def golden_params(fx):
    """Weights of a fixture, re-derived from its seed and checked against the stored checksum."""
    from oracle import hitnet_oracle as O
    sd = O.random_hitnet_params(int(fx["meta"][4]))
    got = np.frombuffer(O.params_digest(sd), dtype=np.uint8)
    assert np.array_equal(got, fx["weights_sha1"]), "RNG drift: fixture weights cannot be re-derived"
    return sd
