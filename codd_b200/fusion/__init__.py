from .fusion import Fusion  # noqa: F401
