import re

import tensorflow as tf


def _snake(name):
    s = re.sub("(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    return re.sub("([a-z])([A-Z])", r"\1_\2", s).lower()


class Layer:
    """[TF] tf.layers.Layer subset: lazy build on first call; variable scope captured on first call
    with default_name = snake_case(class name) (tf/python/layers/base.py _set_scope)."""

    def __init__(self, trainable=True, name=None, dtype=None, **kwargs):
        self.trainable = trainable
        self._base_name = name or _snake(type(self).__name__)
        self._scope = None
        self._dtype = dtype
        self.built = False

    @property
    def dtype(self):
        return self._dtype

    def add_weight(self, name, shape, dtype=None, initializer=None, regularizer=None, trainable=True, constraint=None, **kw):
        return tf.get_variable(name, shape=shape, dtype=dtype or self._dtype or tf.float32, initializer=initializer, trainable=trainable)

    def __call__(self, inputs, *args, **kwargs):
        if self._scope is None:
            with tf.variable_scope(None, default_name=self._base_name) as s:
                self._scope = s
        with tf.variable_scope(self._scope):
            if not self.built:
                if self._dtype is None:
                    self._dtype = inputs.dtype
                self.build(inputs.shape)
                self.built = True
            return self.call(inputs, *args, **kwargs)
