"""Training step of the reference's train.py on the CUDA library: build_model (train.py:35-81), get_optimizer (15-24),
clip_gradients (27-32) and utils.average_gradients (utils.py:34-60).

The reference replicates the graph on `hparams.num_gpus` towers inside one process and averages the tower gradients on a
consolidation device.  Here every tower is one process / one GPU (torchrun); the average is one NCCL all-reduce of the flat
gradient vector.  Everything else -- loss, gradients, global-norm clip, Adam, re-packing of the derived operands -- runs inside
libflowavenet_b200 (fwn_loss_and_grads / fwn_apply_gradients); there is no CPU or autograd fallback.
"""
import ctypes

import torch

from . import _lib


def learning_rate(global_step):
    """train.py:15-20."""
    lr = 0.001
    if not global_step < 200000:
        lr = 0.001 / 2
    if not global_step < 400000:
        lr = 0.001 / 4
    if not global_step < 600000:
        lr = 0.001 / 6
    return lr


def average_flat_gradients(flat, group=None):
    """utils.average_gradients (utils.py:34-60) for towers = processes: in-place mean of the flat gradient vector over `group`
    (one all-reduce; NCCL over NVLink on GPUs, gloo in the CPU tests).  No-op without an initialised process group."""
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return flat
    ws = torch.distributed.get_world_size(group)
    if ws > 1:
        torch.distributed.all_reduce(flat, group=group)
        flat.mul_(1.0 / ws)
    return flat


def broadcast_flat_variables(flat, src=0, group=None):
    """Towers share ONE set of variables in the reference (tf.variable_scope reuse, train.py:51-53).  With one process per tower
    the data-dependent ActNorm initialisation (train.py:221,229) would leave every rank with its own statistics, so rank `src`'s
    variable vector is broadcast after it (documented deviation -- the reference's init step lets the last tower win)."""
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1:
        torch.distributed.broadcast(flat, src=src, group=group)
    return flat


class Trainer:
    """One tower.  `group`: torch.distributed process group over which tower gradients are averaged (None = single tower)."""

    def __init__(self, model, group=None, clip_norm=1.0, beta1=0.9, beta2=0.999, epsilon=1e-8, scale=1.0, split_terms=6):
        if model._precision != _lib.FWN_FP32:
            raise ValueError("training runs on the fp32 engines: use hparams.dtype='float32'")
        self.model, self.group = model, group
        self.clip_norm, self.beta1, self.beta2, self.epsilon = clip_norm, beta1, beta2, epsilon
        self.scale = float(scale)  # hparams.scale (train.py:62,75): static loss scale; 1 here (fp32 accumulation everywhere)
        self.global_step = 0
        L = _lib.lib()
        model._sync_params()
        with torch.cuda.device(model._device):
            _lib.check(L.fwn_train_enable(model._h, _lib.stream_ptr()))
        # split_terms: bf16 products per fp32 product in the training GEMMs.  6 = every product down to 2^-24 (every variable's gradient within
        # 2e-4 of a float64 reference); 3 = ~2^-16..2^-18 per product at half the tensor work (errors up to ~1e-3 of the largest
        # gradient entry on cancellation-heavy gradients; ~10x tighter than bf16).  Inference passes always use 6.
        _lib.check(L.fwn_set_split_terms(model._h, 6, int(split_terms)))
        self._n = L.fwn_grad_floats(model._h)
        self._np = L.fwn_param_floats(model._h)
        self.grads = torch.zeros(self._n, dtype=torch.float32, device=model._device)
        self._ws = None
        self._out = torch.zeros(3, dtype=torch.float32, device=model._device)

    def _workspace(self, B, T):
        need = _lib.lib().fwn_train_workspace_bytes(self.model._h, B, T)
        if need < 0:
            _lib.check(1)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.model._device)
        return self._ws

    def loss_and_grads(self, x, c, g=None):
        """-> (log_p, logdet) device scalars; the flat gradient of -(log_p + logdet) lands in self.grads."""
        m = self.model
        g = m._check_g(g)
        x, c = m._check_xc(x, c, "x")
        B, T = x.shape[0], x.shape[1]
        ws = self._workspace(B, T)
        with torch.cuda.device(m._device):
            _lib.check(_lib.lib().fwn_loss_and_grads(m._h, _lib.ptr(x), _lib.ptr(c), _lib.ptr(g), B, T, _lib.ptr(self._out[0:]),
                                                    _lib.ptr(self._out[1:]), _lib.ptr(self.grads), self._n, _lib.ptr(ws), ws.numel(),
                                                    _lib.stream_ptr()))
        return self._out[0], self._out[1]

    def average_gradients(self):
        """utils.py:34-60 across towers = processes."""
        average_flat_gradients(self.grads[:self._np], self.group)

    def apply_gradients(self):
        """clip_by_global_norm + Adam (train.py:76-81); returns the (pre-clip) global norm as a device scalar."""
        m = self.model
        lr = learning_rate(self.global_step)
        with torch.cuda.device(m._device):
            L = _lib.lib()
            _lib.check(L.fwn_grad_global_norm(m._h, _lib.ptr(self.grads), _lib.ptr(self._out[2:]), _lib.stream_ptr()))
            _lib.check(L.fwn_apply_gradients(m._h, _lib.ptr(self.grads), lr, self.beta1, self.beta2, self.epsilon, self.clip_norm,
                                             self.global_step + 1, _lib.stream_ptr()))
        self.global_step += 1
        return self._out[2], lr

    def train_step(self, x, c, g=None, init=False):
        """One sess.run([..., train_op]) of train.py:221/229/236.  init=True is the ActNorm data-dependent initialisation step."""
        if init:
            self.model.initialize_actnorm(x, c, g)
            self.sync_variables()
        log_p, logdet = self.loss_and_grads(x, c, g)
        self.average_gradients()
        norm, lr = self.apply_gradients()
        return {"log_p": log_p, "logdet": logdet, "loss": -(log_p + logdet), "grad_global_norm": norm, "learning_rate": lr,
                "global_step": self.global_step}

    def sync_variables(self, src=0):
        """All towers adopt rank `src`'s variables (no-op for a single tower)."""
        if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
            return
        if torch.distributed.get_world_size(self.group) < 2:
            return
        m, L = self.model, _lib.lib()
        with torch.cuda.device(m._device):
            flat = torch.empty(self._np, dtype=torch.float32, device=m._device)
            _lib.check(L.fwn_get_train_state(m._h, 0, _lib.ptr(flat), flat.numel(), _lib.stream_ptr()))
            broadcast_flat_variables(flat, src, self.group)
            _lib.check(L.fwn_set_train_state(m._h, 0, _lib.ptr(flat), flat.numel(), _lib.stream_ptr()))

    def state_dict(self):
        """What the reference's Saver writes (train.py:190: variables, Adam slots, global_step), as flat device tensors."""
        m, L = self.model, _lib.lib()
        out = {"global_step": self.global_step}
        with torch.cuda.device(m._device):
            for which, key in enumerate(("variables", "adam_m", "adam_v")):
                t = torch.empty(self._np, dtype=torch.float32, device=m._device)
                _lib.check(L.fwn_get_train_state(m._h, which, _lib.ptr(t), t.numel(), _lib.stream_ptr()))
                out[key] = t
        return out

    def load_state_dict(self, state):
        """Resume (train.py:199-210): variables (every packed operand is re-derived on the device), Adam slots, step counter."""
        m, L = self.model, _lib.lib()
        with torch.cuda.device(m._device):
            for which, key in enumerate(("variables", "adam_m", "adam_v")):
                t = state[key].to(m._device, torch.float32).contiguous()
                _lib.check(L.fwn_set_train_state(m._h, which, _lib.ptr(t), t.numel(), _lib.stream_ptr()))
        self.global_step = int(state["global_step"])

    def gradients(self):
        """{variable name -> view of the flat gradient}."""
        m, L, out = self.model, _lib.lib(), {}
        for i, (k, shp) in enumerate(m.variable_shapes().items()):
            off = L.fwn_param_offset(m._h, i)
            n = 1
            for s in shp:
                n *= s
            out[k] = self.grads[off:off + n].view(shp)
        return out

    def variables(self):
        """{variable name -> fresh copy of the current value} (the handle owns the live variables during training)."""
        m, L, out = self.model, _lib.lib(), {}
        with torch.cuda.device(m._device):
            for k, shp in m.variable_shapes().items():
                t = torch.empty(shp, dtype=torch.float32, device=m._device)
                _lib.check(L.fwn_get_param(m._h, k.encode(), _lib.ptr(t), t.numel(), _lib.stream_ptr()))
                out[k] = t
        return out
