from mmcv.utils import Registry

MODELS = Registry("models")
BACKBONES = MODELS
NECKS = MODELS
HEADS = MODELS
LOSSES = MODELS
SEGMENTORS = MODELS


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_loss(cfg):
    return LOSSES.build(cfg)
