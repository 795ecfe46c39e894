from .backbone import HITUNet  # noqa: F401
from .initialization import TileInitialization  # noqa: F401
from .propagation import TilePropagation  # noqa: F401
from .hitnet import HITNetMF  # noqa: F401
