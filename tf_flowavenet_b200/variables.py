"""Named variables with TF-style scoping (the reference relies on tf.variable_scope / tf.get_variable).

A VariableStore is a dict ``full name -> torch CUDA float32 tensor``.  ``variable_scope`` nests prefixes the way
tf.variable_scope does, so classes built inside each other end up with the reference's names, e.g.
``FloWaveNet/Block_0/Flow_1/AffineCoupling/WaveNet/ResBlock_0_0/Conv_gate/conv1d/kernel``.
"""
import contextlib
import math

_stack = [""]
_default_store = None


class VariableStore(dict):
    def get(self, name, shape=None):
        if name not in self:
            raise KeyError("variable '%s' has not been loaded or initialised" % name)
        v = self[name]
        if shape is not None and tuple(v.shape) != tuple(shape):
            raise ValueError("variable '%s' has shape %s, expected %s" % (name, tuple(v.shape), tuple(shape)))
        return v


def default_store():
    global _default_store
    if _default_store is None:
        _default_store = VariableStore()
    return _default_store


def current_prefix():
    return _stack[-1]


@contextlib.contextmanager
def variable_scope(name, absolute=False):
    full = name if absolute or not _stack[-1] else _stack[-1] + "/" + name
    _stack.append(full)
    try:
        yield full
    finally:
        _stack.pop()


def join(prefix, name):
    return prefix + "/" + name if prefix else name


def he_uniform(shape, rng):
    import numpy as np
    fan_in = int(np.prod(shape[:-1])) if len(shape) > 1 else int(shape[0])
    lim = math.sqrt(6.0 / fan_in)
    return rng.uniform(-lim, lim, shape)
