class SE3:
    def __init__(self, *a, **k):
        raise ImportError("lietorch is not available in this container (import stub)")
