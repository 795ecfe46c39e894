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
from .motion import HRNet, Motion, RAFT3D  # noqa: F401
from .stereo import HITNetMF, HITUNet, TileInitialization, TilePropagation  # noqa: F401

__all__ = ["build_estimator", "Fusion", "Motion", "RAFT3D", "HRNet", "ConsistentOnlineDynamicDepth", "HITNetMF", "HITUNet", "TileInitialization",
           "TilePropagation", "MODELS", "BACKBONES", "ESTIMATORS", "lib", "ops"]


def hitnet_config(max_disp=192):
    """The reference's stereo model dict (configs/models/stereo.py:12-25) without the loss."""
    return dict(type="HITNetMF", backbone=dict(type="HITUNet"),
                initialization=dict(type="TileInitialization", max_disp=max_disp),
                propagation=dict(type="TilePropagation"))


def codd_stereo_config(max_disp=192):
    return dict(type="ConsistentOnlineDynamicDepth", stereo=hitnet_config(max_disp),
                train_cfg=None, test_cfg=dict(mode="whole"))


HRNET_W18_SMALL = dict(   # configs/models/codd.py:48-73
    stage1=dict(num_modules=1, num_branches=1, block="BOTTLENECK", num_blocks=(2,), num_channels=(64,)),
    stage2=dict(num_modules=1, num_branches=2, block="BASIC", num_blocks=(2, 2), num_channels=(18, 36)),
    stage3=dict(num_modules=3, num_branches=3, block="BASIC", num_blocks=(2, 2, 2), num_channels=(18, 36, 72)),
    stage4=dict(num_modules=2, num_branches=4, block="BASIC", num_blocks=(2, 2, 2, 2), num_channels=(18, 36, 72, 144)))


def codd_full_config(max_disp=192, iters=16):
    """The reference's full model dict (configs/models/codd.py:18-101) without the losses."""
    return dict(
        type="ConsistentOnlineDynamicDepth", stereo=hitnet_config(max_disp),
        motion=dict(type="Motion", iters=iters, raft3d=dict(type="RAFT3D", cnet_cfg=dict(
            type="HRNet", norm_cfg=dict(type="SyncBN", requires_grad=False), norm_eval=True, extra=HRNET_W18_SMALL))),
        fusion=dict(type="Fusion", in_channels=24, fusion_channel=32, corr_cfg=dict(type="px2patch", patch_size=3)),
        train_cfg=dict(freeze_stereo=True, freeze_motion=True, freeze_fusion=True), test_cfg=dict(mode="whole"))
