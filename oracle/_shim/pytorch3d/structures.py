class Pointclouds:
    def __init__(self, *a, **k):
        raise ImportError("pytorch3d is not available in this container (import stub)")
