"""TEST STUB of matplotlib.lines."""


class Line2D:
    def __init__(self, *a, **k):
        pass
