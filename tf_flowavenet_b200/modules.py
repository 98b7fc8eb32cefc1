"""WaveNet modules of the coupling network -- same class names, constructor arguments and call signatures as the
reference's modules.py, executing through the per-op C ABI (fp32, channels-last [B, T, C] CUDA tensors).

These classes are the fine-grained surface (unit parity, drop-in use of a single layer).  The throughput path is
``FloWaveNet.forward/reverse`` in model.py, which runs the whole pass in fused kernels.
"""
import torch

from . import _lib
from .variables import VariableStore, current_prefix, default_store, join, variable_scope


def _chk_in(x, C, what):
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 3):
        raise TypeError("%s: expected a float32 CUDA tensor [B, T, C], got %s" % (what, type(x)))
    if x.shape[2] != C:
        raise ValueError("%s: expected %d channels, got %d" % (what, C, x.shape[2]))
    return x.contiguous()


class _Conv1D:
    """Weight-normed Keras Conv1D as built by convolutional.py:53-109 (variables kernel, wn/g, bias)."""

    def __init__(self, prefix, store, in_channels, filters, kernel_size=1, dilation=1, weight_norm=True):
        self.prefix, self.store = prefix, store
        self.cin, self.cout, self.k, self.d, self.wn = in_channels, filters, kernel_size, dilation, weight_norm

    def variables(self):
        names = ["kernel", "bias"] + (["wn/g"] if self.wn else [])
        return {n: join(self.prefix, n) for n in names}

    def __call__(self, x, causal=False, relu=False):
        B, T, _ = x.shape
        w = self.store.get(join(self.prefix, "kernel"), (self.k, self.cin, self.cout))
        b = self.store.get(join(self.prefix, "bias"), (self.cout,))
        g = self.store.get(join(self.prefix, "wn/g"), (self.cout,)) if self.wn else None
        y = torch.empty(B, T, self.cout, device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().fwn_conv1d(_lib.ptr(x), _lib.ptr(w), _lib.ptr(g), _lib.ptr(b), _lib.ptr(y), B, T, self.cin, self.cout,
                                        self.k, self.d, int(causal), int(relu), _lib.stream_ptr()))
        return y


class Conv:
    """modules.py:6-36 -- zero pad + VALID dilated weight-normed conv (causal variant trims the right)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, dilation=1, causal=True, scope='Conv', variables=None):
        self._store = variables if variables is not None else default_store()
        with variable_scope(scope) as vs:
            self._vs = vs
            self._causal = causal
            self._conv = _Conv1D(join(vs, "conv1d"), self._store, in_channels, out_channels, kernel_size, dilation)

    def forward(self, tensor, relu=False):
        x = _chk_in(tensor, self._conv.cin, "Conv")
        return self._conv(x, causal=self._causal, relu=relu)

    def __call__(self, tensor):
        return self.forward(tensor)


class ZeroConv1d:
    """modules.py:39-59 -- 1x1 conv without weight norm, times exp(3 * scale)."""

    def __init__(self, in_channel, out_channel, scope='ZeroConv1d', training_dtype=None, variables=None):
        self._store = variables if variables is not None else default_store()
        with variable_scope(scope) as vs:
            self._vs = vs
        self.cin, self.cout = in_channel, out_channel

    def forward(self, x):
        x = _chk_in(x, self.cin, "ZeroConv1d")
        B, T, _ = x.shape
        w = self._store.get(join(self._vs, "conv1d/kernel"), (1, self.cin, self.cout))
        b = self._store.get(join(self._vs, "conv1d/bias"), (self.cout,))
        s = self._store.get(join(self._vs, "scale"), (1, 1, self.cout))
        y = torch.empty(B, T, self.cout, device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().fwn_zero_conv1d(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(s), _lib.ptr(y), B * T, self.cin, self.cout,
                                             _lib.stream_ptr()))
        return y

    def __call__(self, x):
        return self.forward(x)


def _ew(fn, *tensors, extra=()):
    out = torch.empty_like(tensors[0])
    _lib.check(fn(*[_lib.ptr(t) for t in tensors], _lib.ptr(out), tensors[0].numel(), *extra, _lib.stream_ptr()))
    return out


class ResBlock:
    """modules.py:62-131 -- gated unit with local conditioning, residual and skip 1x1 projections.
    The global-conditioning branch exists in the reference constructor but is never reached."""

    def __init__(self, in_channels, out_channels, skip_channels, kernel_size, dilation, cin_channels=None, local_conditioning=True,
                 global_conditioning=True, causal=False, scope='ResBlock', training_dtype=None, variables=None):
        self._store = variables if variables is not None else default_store()
        with variable_scope(scope) as vs:
            self._vs = vs
            self._causal = causal
            self._local_conditioning = local_conditioning
            self._skip = skip_channels is not None
            self._filter_conv = Conv(in_channels, out_channels, kernel_size, dilation, causal, scope='Conv_filter', variables=self._store)
            self._gate_conv = Conv(in_channels, out_channels, kernel_size, dilation, causal, scope='Conv_gate', variables=self._store)
            # Keras scopes are assigned in first-call order (modules.py:117-127): cond filter, cond gate, res, skip
            n = 0
            if local_conditioning:
                self._filter_conv_c = _Conv1D(join(vs, "conv1d"), self._store, cin_channels, out_channels)
                self._gate_conv_c = _Conv1D(join(vs, "conv1d_1"), self._store, cin_channels, out_channels)
                n = 2
            self._res_conv = _Conv1D(join(vs, "conv1d" + ("_%d" % n if n else "")), self._store, out_channels, out_channels)
            if self._skip:
                self._skip_conv = _Conv1D(join(vs, "conv1d_%d" % (n + 1)), self._store, out_channels, skip_channels)
        self.cin = in_channels

    def forward(self, tensor, c, g=None):
        L = _lib.lib()
        x = _chk_in(tensor, self.cin, "ResBlock")
        h_filter = self._filter_conv(x)
        h_gate = self._gate_conv(x)
        if self._local_conditioning:
            c = _chk_in(c, self._filter_conv_c.cin, "ResBlock(c)")
            h_filter = _ew(L.fwn_add, h_filter, self._filter_conv_c(c), extra=(0,))
            h_gate = _ew(L.fwn_add, h_gate, self._gate_conv_c(c), extra=(0,))
        out = _ew(L.fwn_gated_activation, h_filter, h_gate)
        res = self._res_conv(out)
        skip = self._skip_conv(out) if self._skip else None
        return _ew(L.fwn_residual_scale, x, res), skip

    def __call__(self, tensor, c, g=None):
        return self.forward(tensor, c, g)


class WaveNet:
    """modules.py:134-189.  ``__call__(x, c, g)`` drops g exactly like the reference (modules.py:188-189)."""

    def __init__(self, in_channels=1, out_channels=2, num_blocks=1, num_layers=6, residual_channels=256, gate_channels=256,
                 skip_channels=256, kernel_size=3, cin_channels=80, causal=True, scope='WaveNet', training_dtype=None, variables=None):
        self._store = variables if variables is not None else default_store()
        with variable_scope(scope) as vs:
            self._vs = vs
            self._skip = skip_channels is not None
            self._front_conv = Conv(in_channels, residual_channels, 3, causal=causal, scope='Conv_front', variables=self._store)
            self._res_blocks = []
            for b in range(num_blocks):
                for n in range(num_layers):
                    self._res_blocks.append(ResBlock(residual_channels, gate_channels, skip_channels, kernel_size,
                                                     dilation=kernel_size ** n, cin_channels=cin_channels, causal=causal,
                                                     scope='ResBlock_%d_%d' % (b, n), variables=self._store))
            last = skip_channels if self._skip else residual_channels
            self._final_conv = Conv(last, last, 1, causal=causal, scope='Conv_final', variables=self._store)
            self._final_zero_conv = ZeroConv1d(last, out_channels, variables=self._store)

    def forward(self, x, c, g=None):
        L = _lib.lib()
        h = self._front_conv.forward(x, relu=True)
        skips = []
        for f in self._res_blocks:
            h, s = f(h, c, g)
            skips.append(s)
        if self._skip:  # relu(add_n(skips)) (modules.py:176-177); the relu rides on the last add
            out = skips[0]
            for i, s in enumerate(skips[1:], 1):
                out = _ew(L.fwn_add, out, s, extra=(int(i == len(skips) - 1),))
            if len(skips) == 1:
                out = _ew(L.fwn_add, out, torch.zeros_like(out), extra=(1,))
        else:
            out = _ew(L.fwn_add, h, torch.zeros_like(h), extra=(1,))
        out = self._final_conv.forward(out, relu=True)
        return self._final_zero_conv(out)

    def __call__(self, x, c, g=None):
        return self.forward(x, c)
