import torch.nn as nn

from ..utils import Registry

HOOKS = Registry("hook")


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg

    def init_weights(self):
        pass


def auto_fp16(apply_to=None, out_fp32=False):
    def wrapper(func):
        return func

    return wrapper


class LrUpdaterHook:
    def __init__(self, *args, **kwargs):
        pass
