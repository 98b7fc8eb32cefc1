"""Eager stand-in for the handful of TensorFlow-1.12 symbols the reference's hot path calls.

TEST INFRASTRUCTURE ONLY (lives under oracle/).  Purpose: let the reference's OWN, UNMODIFIED
``model.py`` / ``modules.py`` / ``convolutional.py`` (imported from /root/reference) execute in
this container -- where TF 1.12 cannot be installed -- so that ``tests/golden/make_golden.py``
can produce golden vectors from the reference's composition logic.  Tensors are torch CPU
tensors; "graph construction" simply executes.  Every op below is a restatement of the
published TF semantics ([TF]); nothing here is copied from TensorFlow or from the reference.

Deliberate properties:
  * ``x += y`` never mutates (TF tensors are immutable): see TFTensor.__iadd__.
  * variable_scope / default_name uniquification follows TF's rule (count of opened scopes
    with the same full name), so variable names come out as TF would name them.
  * tf.get_variable always behaves like reuse=AUTO_REUSE (train.py:53 uses it).
  * Variable values can be preset by full name (``set_presets``) so a test controls every weight.
"""
import contextlib
import math

import numpy as _np
import torch as _torch

float16 = _torch.float16
float32 = _torch.float32
float64 = _torch.float64
int32 = _torch.int32
int64 = _torch.int64


class TFTensor(_torch.Tensor):
    """torch tensor with TF's value semantics for augmented assignment."""

    def __iadd__(self, o):
        return self + o

    def __isub__(self, o):
        return self - o

    def __imul__(self, o):
        return self * o

    def get_shape(self):
        return self.shape


class Variable(TFTensor):
    def assign(self, value):
        with _torch.no_grad():
            self.data.copy_(_as_t(value).reshape(self.shape))
        return self


def _as_t(x, dtype=None):
    if not isinstance(x, _torch.Tensor):
        x = _torch.as_tensor(x, dtype=dtype if dtype is not None else (float32 if isinstance(x, float) else None))
    if not isinstance(x, TFTensor):
        x = x.as_subclass(TFTensor)
    return x


def convert_to_tensor(x, dtype=None):
    return _as_t(x, dtype)


constant = convert_to_tensor

# --------------------------------------------------------------------------- scopes
class VariableScope:
    def __init__(self, name):
        self.name = name

    @property
    def original_name_scope(self):
        return self.name + "/" if self.name else ""


_state = {"stack": [VariableScope("")], "opened": {}, "vars": {}, "presets": None, "unused_presets": set(),
          "created_without_preset": [], "rng": _np.random.default_rng(0)}


def reset_default_graph():
    _state.update(stack=[VariableScope("")], opened={}, vars={}, presets=None, unused_presets=set(),
                  created_without_preset=[], rng=_np.random.default_rng(0))


def set_presets(d):
    """Full-name -> array.  Every variable created afterwards must find its value here."""
    _state["presets"] = dict(d)
    _state["unused_presets"] = set(d)


def set_trainable(flag):
    """Shim switch: variables created afterwards take part in tf.gradients (torch autograd leaves)."""
    _state["trainable"] = bool(flag)


def trainable_variables():
    return [v for v in _state["vars"].values() if v.requires_grad]


def gradients(ys, xs, **kw):
    """tf.gradients(ys, xs): symbolic gradient == torch.autograd.grad over the eagerly recorded graph; None where unconnected."""
    return list(_torch.autograd.grad(ys, list(xs), allow_unused=True))


def scalar_mul(scalar, x):
    return x * scalar


def global_variables_dict():
    return dict(_state["vars"])


def shim_report():
    return {"unused_presets": sorted(_state["unused_presets"]), "created_without_preset": list(_state["created_without_preset"])}


def get_variable_scope():
    return _state["stack"][-1]


AUTO_REUSE = "auto_reuse"


@contextlib.contextmanager
def variable_scope(name_or_scope=None, default_name=None, reuse=None, auxiliary_name_scope=True, custom_getter=None, **kw):
    cur = _state["stack"][-1]
    if isinstance(name_or_scope, VariableScope):
        full = name_or_scope.name
    elif name_or_scope is None:
        assert default_name is not None
        base = (cur.name + "/" if cur.name else "") + default_name
        full, idx = base, 0
        while _state["opened"].get(full, 0) > 0:  # [TF] _get_unique_variable_scope
            idx += 1
            full = base + "_%d" % idx
    else:
        full = (cur.name + "/" if cur.name else "") + name_or_scope
    _state["opened"][full] = _state["opened"].get(full, 0) + 1
    vs = VariableScope(full)
    _state["stack"].append(vs)
    try:
        yield vs
    finally:
        _state["stack"].pop()


@contextlib.contextmanager
def name_scope(name=None, *a, **k):
    yield name


@contextlib.contextmanager
def control_dependencies(deps):
    yield


@contextlib.contextmanager
def device(*a, **k):
    yield


def get_variable(name, shape=None, dtype=float32, initializer=None, trainable=True, **kw):
    full = (get_variable_scope().name + "/" if get_variable_scope().name else "") + name
    if full in _state["vars"]:
        return _state["vars"][full]
    shape = tuple(int(s) for s in shape)
    dtype = dtype or float32
    if _state["presets"] is not None and full in _state["presets"]:
        val = _torch.as_tensor(_np.asarray(_state["presets"][full])).to(dtype).reshape(shape)
        _state["unused_presets"].discard(full)
    else:
        if _state["presets"] is not None:
            _state["created_without_preset"].append(full)
        init = initializer if initializer is not None else initializers.glorot_uniform()
        val = init(shape, dtype)
    v = val.clone().as_subclass(Variable)
    if _state.get("trainable") and trainable:
        v.requires_grad_(True)
    _state["vars"][full] = v
    return v


# --------------------------------------------------------------------------- initializers
class _Init:
    def __init__(self, kind, value=0.0):
        self.kind, self.value = kind, value

    def __call__(self, shape, dtype=float32, partition_info=None):
        shape = tuple(int(s) for s in shape)
        if self.kind == "zeros":
            return _torch.zeros(shape, dtype=dtype)
        if self.kind == "const":
            return _torch.full(shape, self.value, dtype=dtype)
        # [TF] variance-scaling fans: receptive field * in / out
        rf = int(_np.prod(shape[:-2])) if len(shape) > 2 else 1
        fan_in = (shape[-2] if len(shape) > 1 else shape[0]) * rf
        fan_out = shape[-1] * rf
        lim = math.sqrt(6.0 / fan_in) if self.kind == "he_uniform" else math.sqrt(6.0 / (fan_in + fan_out))
        return _torch.as_tensor(_state["rng"].uniform(-lim, lim, shape)).to(dtype)


class initializers:  # noqa: N801 (mirrors tf.initializers namespace)
    @staticmethod
    def he_uniform(seed=None):
        return _Init("he_uniform")

    @staticmethod
    def glorot_uniform(seed=None):
        return _Init("glorot_uniform")

    @staticmethod
    def zeros():
        return _Init("zeros")

    @staticmethod
    def constant(value=0.0):
        return _Init("const", value)


# --------------------------------------------------------------------------- ops
def cast(x, dtype, name=None):
    return _as_t(x).to(dtype)


def shape(x):
    return [int(s) for s in x.shape]


def reshape(x, shp):
    return _as_t(x).reshape([int(s) for s in shp])


def transpose(x, perm):
    return _as_t(x).permute(*perm)


def split(value, num_or_size_splits, axis=0):
    n = value.shape[axis] // num_or_size_splits
    return list(_torch.split(_as_t(value), n, dim=axis))


def concat(values, axis):
    return _torch.cat([_as_t(v) for v in values], dim=axis)


def pad(tensor, paddings):
    flat = []
    for lo, hi in reversed(list(paddings)):
        flat += [int(lo), int(hi)]
    return _torch.nn.functional.pad(_as_t(tensor), flat)


def reduce_mean(x, axis=None, keepdims=False):
    x = _as_t(x)
    if axis is None:
        return x.mean()
    return x.mean(dim=tuple(axis) if isinstance(axis, (list, tuple)) else axis, keepdim=keepdims)


def reduce_sum(x, axis=None, keepdims=False):
    x = _as_t(x)
    if axis is None:
        return x.sum()
    return x.sum(dim=tuple(axis) if isinstance(axis, (list, tuple)) else axis, keepdim=keepdims)


def exp(x):
    return _torch.exp(_as_t(x))


def log(x):
    return _torch.log(_as_t(x))


def sqrt(x):
    return _torch.sqrt(_as_t(x))


def pow(x, y):  # noqa: A001
    return _torch.pow(_as_t(x), y)


def tanh(x):
    return _torch.tanh(_as_t(x))


def sigmoid(x):
    return _torch.sigmoid(_as_t(x))


def add_n(xs):
    out = xs[0]
    for t in xs[1:]:
        out = out + t
    return out


def expand_dims(x, axis):
    return _as_t(x).unsqueeze(axis)


def squeeze(x, axis=None):
    return _as_t(x).squeeze() if axis is None else _as_t(x).squeeze(axis)


def tile(x, multiples):
    return _as_t(x).repeat(*[int(m) for m in multiples])


def cond(pred, true_fn, false_fn):
    return true_fn() if bool(pred) else false_fn()


def random_normal(shp, dtype=float32):
    return _as_t(_torch.as_tensor(_state["rng"].standard_normal([int(s) for s in shp])).to(dtype))


class nn:  # noqa: N801
    @staticmethod
    def relu(x):
        return _torch.relu(_as_t(x))

    @staticmethod
    def leaky_relu(x, alpha=0.2):
        x = _as_t(x)
        return _torch.maximum(alpha * x, x)

    @staticmethod
    def embedding_lookup(params, ids):
        return _as_t(params)[_torch.as_tensor(ids).long()]


class _HParams:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def values(self):
        return dict(self.__dict__)


class _Training:
    HParams = _HParams


class contrib:  # noqa: N801
    training = _Training
