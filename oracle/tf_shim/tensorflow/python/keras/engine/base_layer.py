class InputSpec:
    def __init__(self, ndim=None, axes=None, **kw):
        self.ndim, self.axes = ndim, axes
