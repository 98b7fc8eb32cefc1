"""FloWaveNet flow model -- same class names, constructor arguments and call signatures as the reference's model.py.

``FloWaveNet.forward / reverse / upsample`` run the fused whole-pass kernels behind one C-ABI handle (fp32 parity mode
or bf16/tcgen05 mixed mode, chosen by ``hparams.dtype``).  ``Block / Flow / AffineCoupling / ActNorm / change_order``
execute one reference op at a time through the per-op C ABI and are what the unit parity tests exercise.
All tensors are float32 CUDA tensors in the reference's channels-last layout [B, T, C].
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .modules import WaveNet, _chk_in
from .variables import VariableStore, default_store, he_uniform, join, variable_scope


def _rows(x):
    return x.shape[0] * x.shape[1]


class ActNorm:
    """model.py:7-105.  forward: y=(x+b)exp(3 logs), objective=mean_c(3 logs); reverse: x=y exp(-3 logs)-b.
    ``init=True`` performs the data-dependent initialisation on the next forward call (the reference needs
    a *tensor* init for that -- a Python bool is a no-op there; here True means 'do it')."""

    def __init__(self, in_channel, logdet=True, init=False, logscale=3., scope='ActNorm', training_dtype=None, variables=None):
        if logscale != 3.:
            raise NotImplementedError("logscale is fixed to 3 as in every reference call site")
        self._store = variables if variables is not None else default_store()
        with variable_scope(scope) as vs:
            self._vs = vs
        self._in_channel, self._logdet, self._init = in_channel, logdet, init

    def _vars(self):
        C = self._in_channel
        return self._store.get(join(self._vs, "b"), (1, 1, C)), self._store.get(join(self._vs, "logs"), (1, 1, C))

    def forward(self, x):
        x = _chk_in(x, self._in_channel, "ActNorm")
        L, C = _lib.lib(), self._in_channel
        if self._init:
            for n in ("b", "logs"):
                self._store.setdefault(join(self._vs, n), torch.zeros(1, 1, C, device=x.device))
            b, logs = self._vars()
            _lib.check(L.fwn_actnorm_ddi(_lib.ptr(x), _lib.ptr(b), _lib.ptr(logs), _rows(x), C, _lib.stream_ptr()))
            self._init = False
        b, logs = self._vars()
        y = torch.empty_like(x)
        obj = torch.empty((), device=x.device, dtype=torch.float32)
        _lib.check(L.fwn_actnorm_fwd(_lib.ptr(x), _lib.ptr(b), _lib.ptr(logs), _lib.ptr(y), _lib.ptr(obj), _rows(x), C, _lib.stream_ptr()))
        return (y, obj) if self._logdet else y

    def reverse(self, x):
        x = _chk_in(x, self._in_channel, "ActNorm")
        b, logs = self._vars()
        y = torch.empty_like(x)
        _lib.check(_lib.lib().fwn_actnorm_rev(_lib.ptr(x), _lib.ptr(b), _lib.ptr(logs), _lib.ptr(y), _rows(x), self._in_channel,
                                             _lib.stream_ptr()))
        return y

    def __call__(self, x):
        return self.forward(x)


class AffineCoupling:
    """model.py:108-164."""

    def __init__(self, in_channel, cin_channel, filter_size=256, num_layer=6, affine=True, causal=False, scope='AffineCoupling',
                 training_dtype=None, variables=None):
        self._store = variables if variables is not None else default_store()
        with variable_scope(scope) as vs:
            self._vs = vs
            self._in_channel, self._affine = in_channel, affine
            self._net = WaveNet(in_channels=in_channel // 2, out_channels=in_channel if affine else in_channel // 2, num_blocks=1,
                                num_layers=num_layer, residual_channels=filter_size, gate_channels=filter_size,
                                skip_channels=filter_size, kernel_size=3, cin_channels=cin_channel // 2, causal=causal,
                                variables=self._store)

    def _net_out(self, x, c, g):
        h = self._in_channel // 2
        in_a = x[:, :, :h].contiguous()           # tf.split (model.py:124-125)
        c_a = c[:, :, :c.shape[2] // 2].contiguous()
        return self._net(in_a, c_a, None if g is None else g[:, :, :g.shape[2] // 2])

    def forward(self, x, c, g=None):
        x = _chk_in(x, self._in_channel, "AffineCoupling")
        net = self._net_out(x, c, g)
        y = torch.empty_like(x)
        ld = torch.empty((), device=x.device, dtype=torch.float32) if self._affine else None
        _lib.check(_lib.lib().fwn_affine_fwd(_lib.ptr(x), _lib.ptr(net), _lib.ptr(y), _lib.ptr(ld), _rows(x), self._in_channel,
                                            int(self._affine), _lib.stream_ptr()))
        return y, ld

    def reverse(self, output, c, g=None):
        x = _chk_in(output, self._in_channel, "AffineCoupling")
        net = self._net_out(x, c, g)
        y = torch.empty_like(x)
        _lib.check(_lib.lib().fwn_affine_rev(_lib.ptr(x), _lib.ptr(net), _lib.ptr(y), _rows(x), self._in_channel, int(self._affine),
                                            _lib.stream_ptr()))
        return y

    def __call__(self, x, c, g=None):
        return self.forward(x, c, g)


def _swap(t):
    t = t.contiguous()
    y = torch.empty_like(t)
    _lib.check(_lib.lib().fwn_change_order(_lib.ptr(t), _lib.ptr(y), _rows(t), t.shape[2], _lib.stream_ptr()))
    return y


def change_order(x, c, g=None):
    """model.py:166-174: swap the channel halves of x, c (and g)."""
    return _swap(x), _swap(c), (None if g is None else _swap(g))


class Flow:
    """model.py:176-205."""

    def __init__(self, in_channel, cin_channel, filter_size, num_layer, init, affine=True, causal=False, scope='Flow',
                 training_dtype=None, variables=None):
        store = variables if variables is not None else default_store()
        with variable_scope(scope) as vs:
            self._vs = vs
            self._actnorm = ActNorm(in_channel, init=init, variables=store)
            self._coupling = AffineCoupling(in_channel, cin_channel, filter_size=filter_size, num_layer=num_layer, affine=affine,
                                            causal=causal, variables=store)

    def forward(self, x, c, g=None):
        out, logdet = self._actnorm(x)
        out, det = self._coupling(out, c, g)
        out, c, g = change_order(out, c, g)
        if det is not None:
            logdet = logdet + det
        return out, c, g, logdet

    def reverse(self, output, c, g=None):
        output, c, g = change_order(output, c, g)
        x = self._coupling.reverse(output, c, g)
        return self._actnorm.reverse(x), c, g

    def __call__(self, x, c, g=None):
        return self.forward(x, c, g)


def _squeeze(t, inverse=False):
    t = t.contiguous()
    B, T, C = t.shape
    L = _lib.lib()
    if not inverse:
        y = torch.empty(B, T // 2, 2 * C, device=t.device, dtype=torch.float32)
        _lib.check(L.fwn_squeeze(_lib.ptr(t), _lib.ptr(y), B, T, C, _lib.stream_ptr()))
    else:
        y = torch.empty(B, T * 2, C // 2, device=t.device, dtype=torch.float32)
        _lib.check(L.fwn_unsqueeze(_lib.ptr(t), _lib.ptr(y), B, T, C, _lib.stream_ptr()))
    return y


class Block:
    """model.py:207-280: squeeze x, c (, g) then n_flow flows; reverse undoes them and unsqueezes."""

    def __init__(self, in_channel, cin_channel, n_flow, n_layer, init, affine=True, causal=False, scope='Block', training_dtype=None,
                 variables=None):
        store = variables if variables is not None else default_store()
        with variable_scope(scope) as vs:
            self._vs = vs
            self._flows = [Flow(in_channel * 2, cin_channel * 2, init=init, filter_size=256, num_layer=n_layer, affine=affine,
                                causal=causal, scope='Flow_%d' % i, variables=store) for i in range(n_flow)]

    def forward(self, x, c, g=None):
        out, c = _squeeze(x), _squeeze(c)
        g = None if g is None else _squeeze(g)
        logdet = None
        for flow in self._flows:
            out, c, g, det = flow(out, c, g)
            logdet = det if logdet is None else logdet + det
        return out, c, g, logdet

    def reverse(self, output, c, g=None):
        x = output
        for flow in self._flows[::-1]:
            x, c, g = flow.reverse(x, c, g)
        return _squeeze(x, True), _squeeze(c, True), (None if g is None else _squeeze(g, True))

    def __call__(self, x, c, g=None):
        return self.forward(x, c, g)


# hparams.dtype -> numeric mode.  'float16' is the reference's own mixed dtype (hparams.py:9: fp16 compute, fp32 master weights);
# 'bfloat16' / 'mixed' keep bf16 operands (the round-1 throughput mode; same speed, 8x coarser operand rounding).
_PRECISION = {"float32": _lib.FWN_FP32, "fp32": _lib.FWN_FP32, "bfloat16": _lib.FWN_MIXED_BF16, "bf16": _lib.FWN_MIXED_BF16,
              "float16": _lib.FWN_MIXED_FP16, "fp16": _lib.FWN_MIXED_FP16, "half": _lib.FWN_MIXED_FP16, "mixed": _lib.FWN_MIXED_BF16}


class FloWaveNet:
    """model.py:282-404.

    forward(x, c, g=None) -> (log_p, logdet)  [two float32 device scalars, like the reference]
    reverse(z, c, g=None) -> x [B, T, 1]
    upsample(c)           -> [B, T, num_mels]
    Extra, non-reference accessors: ``forward(..., return_z=True)``, ``load_variables``, ``init_variables``,
    ``variables()``, ``initialize_actnorm(x, c)`` (the DDI step of train.py:221,229).
    """

    def __init__(self, hparams, init=False, scope='FloWaveNet', variables=None, device=None):
        self._hparams = hparams
        self._store = variables if variables is not None else default_store()
        self._device = torch.device(device if device is not None else "cuda")
        self._init = init
        dt = hparams.dtype if isinstance(hparams.dtype, str) else str(hparams.dtype).split(".")[-1]
        if dt not in _PRECISION:
            raise ValueError("unsupported hparams.dtype %r" % (hparams.dtype,))
        self._precision = _PRECISION[dt]
        self._n_block, self._cin_channels = hparams.n_block, hparams.num_mels
        cfg = _lib.FwnConfig()
        cfg.n_block, cfg.n_flow, cfg.n_layer, cfg.num_mels = hparams.n_block, hparams.n_flow, hparams.n_layer, hparams.num_mels
        cfg.filter_size, cfg.affine, cfg.causal = 256, int(hparams.affine), int(hparams.causality)  # Block hard-codes 256
        cfg.n_upsample = len(hparams.upsample_scales)
        for i, s in enumerate(hparams.upsample_scales):
            cfg.upsample_scales[i] = int(s)
        cfg.gin_channels, cfg.n_speakers, cfg.precision = int(hparams.gin_channels), int(hparams.n_speakers), self._precision
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().fwn_create(ctypes.byref(cfg), ctypes.byref(self._h)))
        self._device = torch.device("cuda", torch.cuda.current_device()) if self._device.index is None else self._device
        self._hop = int(np.prod(hparams.upsample_scales))
        self._dirty = True
        self._ws = None
        with variable_scope(scope) as vs:
            self._vs = vs
            self._blocks = []
            in_ch, cin_ch = 1, self._cin_channels
            for i in range(self._n_block):  # per-op mirror of the graph (model.py:295-299), shares the variable store
                self._blocks.append(Block(in_ch, cin_ch, hparams.n_flow, hparams.n_layer, init=init, affine=hparams.affine,
                                          causal=hparams.causality, scope='Block_%d' % i, variables=self._store))
                in_ch, cin_ch = in_ch * 2, cin_ch * 2

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                _lib.lib().fwn_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- variables -----------------------------------------------------------------------------------------------
    def variable_shapes(self):
        L, out = _lib.lib(), {}
        name, shape, rank = ctypes.c_char_p(), (ctypes.c_int64 * 4)(), ctypes.c_int()
        for i in range(L.fwn_num_params(self._h)):
            _lib.check(L.fwn_param_info(self._h, i, ctypes.byref(name), ctypes.byref(shape), ctypes.byref(rank)))
            out[name.value.decode()] = tuple(shape[j] for j in range(rank.value))
        return out

    def load_variables(self, values):
        """values: {name relative to the model scope -> array-like}; e.g. a converted TF checkpoint."""
        shapes = self.variable_shapes()
        self._refresh_store()   # the handle may own newer values (DDI, optimizer steps): keep them for the names not in `values`
        for k, v in values.items():
            if k not in shapes:
                raise KeyError("unknown variable '%s'" % k)
            t = torch.as_tensor(np.asarray(v), dtype=torch.float32).reshape(shapes[k]).to(self._device).contiguous()
            self._store[join(self._vs, k)] = t
        self._dirty = True

    def init_variables(self, seed=0, zero_init_coupling=True):
        """Initialisers of the reference: he-uniform kernels/biases (modules.py:21-22,79-97), wn/g = 1
        (convolutional.py:77), zero ZeroConv1d (modules.py:46-49), ActNorm b = logs = 0 (to be set by DDI)."""
        rng = np.random.default_rng(seed)
        vals = {}
        for k, shp in self.variable_shapes().items():
            if k.endswith("wn/g"):
                a = np.ones(shp)
            elif "ZeroConv1d" in k or "/ActNorm/" in k or (k.startswith("conv2d_transpose") and k.endswith("bias")):
                a = np.zeros(shp) if (zero_init_coupling or "ZeroConv1d" not in k) else rng.uniform(-0.02, 0.02, shp)
            elif k.endswith("bias"):
                a = rng.uniform(-1, 1, shp) * np.sqrt(6.0 / shp[0])  # he_uniform on a 1-D shape: fan_in = shape[0]
            elif k == "speaker_embeddings":
                a = rng.standard_normal(shp) * 0.1
            else:
                a = he_uniform(shp if len(shp) == 3 else (shp[0] * shp[1], 1), rng).reshape(shp)
            vals[k] = a
        self.load_variables(vals)

    def _refresh_store(self):
        """Copy the handle's live variables back into the store.  Once the variables have been uploaded the HANDLE owns them: the
        data-dependent ActNorm initialisation and every optimizer step (Trainer) update them on the device, so the store is stale
        until refreshed -- reading or partially overwriting it without this would mix trained and initial values."""
        if self._dirty or not getattr(self, "_h", None):
            return
        L = _lib.lib()
        with torch.cuda.device(self._device):
            for k, shp in self.variable_shapes().items():
                name = join(self._vs, k)
                t = self._store.get(name) if name in self._store else None
                if t is None or tuple(t.shape) != tuple(shp) or not t.is_contiguous() or t.device != self._device:
                    t = torch.empty(shp, dtype=torch.float32, device=self._device)
                    self._store[name] = t
                _lib.check(L.fwn_get_param(self._h, k.encode(), _lib.ptr(t), t.numel(), _lib.stream_ptr()))

    def variables(self):
        """{relative name -> CUDA tensor} of the CURRENT values (what tf.train.Saver would write, train.py:190,252): refreshed from
        the handle, which owns the live variables after DDI / training steps."""
        self._refresh_store()
        return {k: self._store.get(join(self._vs, k)) for k in self.variable_shapes()}

    def _sync_params(self):
        if not self._dirty:
            return
        L = _lib.lib()
        with torch.cuda.device(self._device):
            for k, shp in self.variable_shapes().items():
                t = self._store.get(join(self._vs, k), shp)
                _lib.check(L.fwn_set_param(self._h, k.encode(), _lib.ptr(t), t.numel(), _lib.stream_ptr()))
            _lib.check(L.fwn_prepack(self._h, _lib.stream_ptr()))
        self._dirty = False

    def _workspace(self, B, T):
        need = _lib.lib().fwn_workspace_bytes(self._h, B, T)
        if need < 0:
            _lib.check(1)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self._device)
        return self._ws

    def _check_xc(self, x, c, what):
        for t, n in ((x, what), (c, "c")):
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dim() == 3):
                raise TypeError("%s must be a CUDA tensor of rank 3" % n)
        if x.shape[2] != 1 or c.shape[2] != self._cin_channels or c.shape[0] != x.shape[0]:
            raise ValueError("expected %s [B,T,1] and c [B,T/hop,%d]" % (what, self._cin_channels))
        if c.shape[1] * self._hop != x.shape[1]:
            raise ValueError("len(%s)=%d must equal len(c)*hop=%d*%d (tfrecord.py:53)" % (what, x.shape[1], c.shape[1], self._hop))
        return x.float().contiguous(), c.float().contiguous()  # tf.cast to hparams.dtype (model.py:323-324)

    def _check_g(self, g):
        if g is None and self._hparams.gin_channels > 0:
            raise ValueError('g is None')  # model.py:320-321, 353-354
        return None if g is None else g.to(self._device, torch.int32).contiguous()

    # ---- the path ------------------------------------------------------------------------------------------------
    def forward(self, x, c, g=None, return_z=False, _ddi=False):
        g = self._check_g(g)
        x, c = self._check_xc(x, c, "x")
        self._sync_params()
        B, T = x.shape[0], x.shape[1]
        ws = self._workspace(B, T)
        out = torch.empty(2, device=self._device, dtype=torch.float32)
        z = torch.empty_like(x) if return_z else None
        ddi = bool(_ddi or self._init)
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().fwn_forward(self._h, _lib.ptr(x), _lib.ptr(c), _lib.ptr(g), B, T, _lib.ptr(z), _lib.ptr(out[0:]),
                                             _lib.ptr(out[1:]), int(ddi), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
        self._init = False
        return (out[0], out[1], z) if return_z else (out[0], out[1])

    def initialize_actnorm(self, x, c, g=None):
        """The `sess.run(..., feed_dict={init: True})` step of train.py:221,229."""
        return self.forward(x, c, g, _ddi=True)

    def reverse(self, z, c, g=None):
        g = self._check_g(g)
        z, c = self._check_xc(z, c, "z")
        self._sync_params()
        B, T = z.shape[0], z.shape[1]
        ws = self._workspace(B, T)
        x = torch.empty_like(z)
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().fwn_reverse(self._h, _lib.ptr(z), _lib.ptr(c), _lib.ptr(g), B, T, _lib.ptr(x), _lib.ptr(ws), ws.numel(),
                                             _lib.stream_ptr()))
        return x

    def upsample(self, c):
        """model.py:398-404 through the per-op entry point (reference layout in, reference layout out)."""
        c = c.float().contiguous()
        L = _lib.lib()
        for i, s in enumerate(self._hparams.upsample_scales):
            n = "conv2d_transpose" if i == 0 else "conv2d_transpose_%d" % i
            k, g, b = (self._store.get(join(self._vs, n + "/" + v)) for v in ("kernel", "wn/g", "bias"))
            B, Tm, M = c.shape
            y = torch.empty(B, Tm * s, M, device=c.device, dtype=torch.float32)
            _lib.check(L.fwn_upsample_stage(_lib.ptr(c), _lib.ptr(k), _lib.ptr(g), _lib.ptr(b), _lib.ptr(y), B, Tm, M, int(s),
                                           _lib.stream_ptr()))
            c = y
        return c

    # per-op (unfused) execution of the same graph, mirroring model.py:338-347 / 374-396 line by line
    def forward_unfused(self, x, c, g=None):
        self._check_g(g)
        x, c = self._check_xc(x, c, "x")
        out, c, logdet = x, self.upsample(c), None
        for block in self._blocks:
            out, c, _, det = block(out, c, None)
            logdet = det if logdet is None else logdet + det
        lp = torch.empty((), device=x.device, dtype=torch.float32)
        oc = out.contiguous()
        _lib.check(_lib.lib().fwn_log_p(_lib.ptr(oc), _lib.ptr(lp), oc.numel(), _lib.stream_ptr()))
        return lp, logdet

    def reverse_unfused(self, z, c, g=None):
        self._check_g(g)
        x, c = self._check_xc(z, c, "z")
        c = self.upsample(c)
        for _ in range(self._n_block):
            x, c = _squeeze(x), _squeeze(c)
        for block in self._blocks[::-1]:
            x, c, _ = block.reverse(x, c, None)
        return x

    def reverse_chunk(self, z_ext, c_ext, halo_l, halo_r):
        """Inverse pass of a time chunk extended by real neighbour data (fwn_reverse_chunk); returns the interior."""
        z, c = self._check_xc(z_ext, c_ext, "z")
        self._sync_params()
        B, Te = z.shape[0], z.shape[1]
        ws = self._workspace(B, Te)
        x = torch.empty(B, Te - halo_l - halo_r, 1, device=self._device, dtype=torch.float32)
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().fwn_reverse_chunk(self._h, _lib.ptr(z), _lib.ptr(c), B, Te, int(halo_l), int(halo_r), _lib.ptr(x),
                                                   _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
        return x

    def reverse_sharded(self, z_local, c_local, rank, world, group=None):
        """Time-chunk sharded synthesis of one long utterance (see sharding.py)."""
        from . import sharding
        return sharding.reverse_sharded(self.reverse_chunk, self._hparams, z_local, c_local, rank, world, group)

    def receptive_halo(self):
        return _lib.lib().fwn_receptive_halo(self._h)

    PROFILE_FAMILIES = ("front_conv", "gate_gemm", "res_skip_gemm", "final_conv", "zero_affine", "upsample")

    def profile(self, on=True):
        _lib.check(_lib.lib().fwn_profile_enable(self._h, int(on)))

    def profile_read(self):
        """{family: (ms, launches, algorithmic work)} accumulated since the last read (CUDA events on the pass's stream)."""
        ms, n, w = (ctypes.c_double * 8)(), (ctypes.c_int64 * 8)(), (ctypes.c_double * 8)()
        _lib.check(_lib.lib().fwn_profile_read(self._h, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(w)))
        return {k: (ms[i], n[i], w[i]) for i, k in enumerate(self.PROFILE_FAMILIES)}

    def last_launches(self):
        return _lib.lib().fwn_last_launches(self._h)

    def set_layer_fusion(self, mode):
        """Mixed-precision passes: 1 = fused ResBlock-layer kernel, 0 = gate GEMM + res|skip GEMM as two launches, -1 = default."""
        _lib.check(_lib.lib().fwn_set_layer_fusion(self._h, int(mode)))

    # host-buffer entry points (numpy / pinned torch CPU tensors): the end-to-end call of synthesize.py:44-46
    def _check_host(self, a, c, g, what):
        """Same validation as _check_xc / _check_g, for HOST buffers: float32, contiguous, [B,T,1] / [B,T/hop,mels] / int32 [B]."""
        a, c = (t if isinstance(t, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32)) for t in (a, c))
        for t, n in ((a, what), (c, "c")):
            if t.is_cuda or t.dim() != 3:
                raise TypeError("%s must be a host (CPU) array of rank 3" % n)
        if a.shape[2] != 1 or c.shape[2] != self._cin_channels or c.shape[0] != a.shape[0]:
            raise ValueError("expected %s [B,T,1] and c [B,T/hop,%d]" % (what, self._cin_channels))
        if c.shape[1] * self._hop != a.shape[1]:
            raise ValueError("len(%s)=%d must equal len(c)*hop=%d*%d (tfrecord.py:53)" % (what, a.shape[1], c.shape[1], self._hop))
        if g is None and self._hparams.gin_channels > 0:
            raise ValueError('g is None')  # model.py:320-321, 353-354
        if g is not None:
            g = (g if isinstance(g, torch.Tensor) else torch.from_numpy(np.asarray(g))).to("cpu", torch.int32).contiguous()
            if g.dim() != 1 or g.shape[0] != a.shape[0]:
                raise ValueError("g must hold one speaker id per utterance")
        return a.float().contiguous(), c.float().contiguous(), g

    def reverse_host(self, z, c, out=None, g=None):
        self._sync_params()
        z, c, g = self._check_host(z, c, g, "z")
        B, T = z.shape[0], z.shape[1]
        out = torch.empty(B, T, 1, dtype=torch.float32, pin_memory=True) if out is None else out
        if tuple(out.shape) != (B, T, 1) or out.dtype != torch.float32 or out.is_cuda or not out.is_contiguous():
            raise ValueError("out must be a contiguous float32 host tensor of shape [B,T,1]")
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().fwn_reverse_host(self._h, _lib.ptr(z), _lib.ptr(c), _lib.ptr(g), B, T, _lib.ptr(out)))
        return out

    def forward_host(self, x, c, z_out=None, g=None):
        self._sync_params()
        x, c, g = self._check_host(x, c, g, "x")
        B, T = x.shape[0], x.shape[1]
        lp, ld = ctypes.c_float(), ctypes.c_float()
        with torch.cuda.device(self._device):
            _lib.check(_lib.lib().fwn_forward_host(self._h, _lib.ptr(x), _lib.ptr(c), _lib.ptr(g), B, T, _lib.ptr(z_out), ctypes.byref(lp),
                                                  ctypes.byref(ld)))
        return lp.value, ld.value
