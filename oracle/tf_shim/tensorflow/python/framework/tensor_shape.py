"""[TF] TensorShape / Dimension subset used by convolutional.py:54-62,156-167."""


class Dimension:
    def __init__(self, v):
        self.value = None if v is None else int(v)

    def __int__(self):
        return self.value

    __index__ = __int__


class TensorShape:
    def __init__(self, dims):
        if isinstance(dims, TensorShape):
            dims = [d.value for d in dims.dims]
        self.dims = [d if isinstance(d, Dimension) else Dimension(d) for d in dims]

    def __len__(self):
        return len(self.dims)

    def __getitem__(self, i):
        return self.dims[i]

    def as_list(self):
        return [d.value for d in self.dims]
