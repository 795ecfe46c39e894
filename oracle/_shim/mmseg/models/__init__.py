from . import builder  # noqa: F401
from .builder import BACKBONES, LOSSES, MODELS  # noqa: F401
