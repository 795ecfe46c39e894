from .extractor import BasicEncoder  # noqa: F401
from .hrnet import HRNet  # noqa: F401
from .motion import Motion  # noqa: F401
from .raft3d import RAFT3D, BasicUpdateBlock, ConvGRU, ResizeConcatConv  # noqa: F401
