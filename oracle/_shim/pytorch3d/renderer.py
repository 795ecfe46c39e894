class _Absent:
    def __init__(self, *a, **k):
        raise ImportError("pytorch3d is not available in this container (import stub)")


class PerspectiveCameras(_Absent):
    pass


class PointsRasterizationSettings(_Absent):
    pass


class PointsRenderer:  # subclassed by the reference at import time
    def __init__(self, *a, **k):
        raise ImportError("pytorch3d is not available in this container (import stub)")


class PointsRasterizer(_Absent):
    pass


class AlphaCompositor(_Absent):
    pass
