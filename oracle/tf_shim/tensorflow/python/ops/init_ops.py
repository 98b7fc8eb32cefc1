import tensorflow as tf


def zeros_initializer():
    return tf.initializers.zeros()


def constant_initializer(value=0.0):
    return tf.initializers.constant(value)
