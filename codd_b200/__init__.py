"""codd_b200 — B200-native (sm_100a) implementation of CODD's per-frame stereo hot path.

The package holds the CUDA kernels + C ABI (``csrc/``, ``libcodd_b200.so``) and the host-side
mirror of the reference's plugin interface: registry-built ``nn.Module`` classes with the
reference's names, constructor arguments, method contracts and ``state_dict`` keys.
"""
from . import lib, ops  # noqa: F401
from .builder import build_estimator  # noqa: F401
from .codd import ConsistentOnlineDynamicDepth  # noqa: F401
from .registry import BACKBONES, ESTIMATORS, MODELS  # noqa: F401
from .fusion import Fusion  # noqa: F401
from .stereo import HITNetMF, HITUNet, TileInitialization, TilePropagation  # noqa: F401

__all__ = ["build_estimator", "Fusion", "ConsistentOnlineDynamicDepth", "HITNetMF", "HITUNet", "TileInitialization",
           "TilePropagation", "MODELS", "BACKBONES", "ESTIMATORS", "lib", "ops"]


def hitnet_config(max_disp=192):
    """The reference's stereo model dict (configs/models/stereo.py:12-25) without the loss."""
    return dict(type="HITNetMF", backbone=dict(type="HITUNet"),
                initialization=dict(type="TileInitialization", max_disp=max_disp),
                propagation=dict(type="TilePropagation"))


def codd_stereo_config(max_disp=192):
    return dict(type="ConsistentOnlineDynamicDepth", stereo=hitnet_config(max_disp),
                train_cfg=None, test_cfg=dict(mode="whole"))
