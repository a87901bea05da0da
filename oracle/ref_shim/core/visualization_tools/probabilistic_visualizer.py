class ProbabilisticVisualizer:  # matplotlib-based debug GUI in the reference; not on the path
    def __init__(self, *a, **k):
        raise NotImplementedError
