"""Training step of the reference's train.py on the CUDA library: build_model (train.py:35-81), get_optimizer (15-24),
clip_gradients (27-32) and utils.average_gradients (utils.py:34-60).

The reference replicates the graph on `hparams.num_gpus` towers inside one process and averages the tower gradients on a
consolidation device.  Here every tower is one process / one GPU (torchrun); the average is an NCCL all-reduce of the flat
gradient vector, issued per block-sized bucket in the order the backward pass finishes them (last block first) on a
communication stream, so it overlaps the rest of the backward pass.  Everything else -- loss, gradients, global-norm clip, Adam, re-packing of the derived operands -- runs inside
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


def bucket_ranges(variable_shapes, n_block):
    """[(offset, count)] of the gradient buckets in production order -- the Python twin of fwn_grad_bucket_range: block n-1 first,
    ..., block 0, then the upsampler variables, then the speaker embeddings if present.  `variable_shapes`: ordered
    {name -> shape} as FloWaveNet.variable_shapes() returns it; every variable occupies ceil4(numel) floats of the flat vector."""
    off, starts, total = 0, {}, 0
    for k, shp in variable_shapes.items():
        n = 1
        for s in shp:
            n *= int(s)
        starts[k] = off
        off += (n + 3) & ~3
    total = off
    blk = [starts["Block_%d/Flow_0/ActNorm/b" % i] for i in range(n_block)]
    blk.append(starts.get("speaker_embeddings", total))
    out = [(blk[i], blk[i + 1] - blk[i]) for i in range(n_block - 1, -1, -1)]
    out.append((0, blk[0]))
    if blk[n_block] < total:
        out.append((blk[n_block], total - blk[n_block]))
    return out


def average_flat_gradients_bucketed(flat, buckets, group=None, waiters=None, stream=None):
    """utils.average_gradients (utils.py:34-60), one all-reduce per bucket in production order.  With `waiters` (one callable per
    bucket, each makes the current CUDA stream wait until that bucket is final) and a communication `stream`, the reductions are
    enqueued asynchronously behind those waits and overlap whatever the producer is still computing; returns the pending work
    handles (call .wait() on each before reading `flat`).  Without them (CPU / gloo) the buckets are reduced in place, blocking."""
    dist = torch.distributed
    if not (dist.is_available() and dist.is_initialized()):
        return []
    ws = dist.get_world_size(group)
    if ws < 2:
        return []
    nccl = dist.get_backend(group) == "nccl"
    op = dist.ReduceOp.AVG if nccl else dist.ReduceOp.SUM   # gloo has no AVG: sum, then scale
    pending = []
    for k, (off, cnt) in enumerate(buckets):
        if cnt == 0:
            continue
        piece = flat[off:off + cnt]
        if stream is not None:
            with torch.cuda.stream(stream):
                if waiters is not None:
                    waiters[k]()
                pending.append(dist.all_reduce(piece, op=op, group=group, async_op=True))
        else:
            dist.all_reduce(piece, op=op, group=group)
            if not nccl:
                piece.mul_(1.0 / ws)
    return pending


def broadcast_flat_variables(flat, src=0, group=None):
    """Towers share ONE set of variables in the reference (tf.variable_scope reuse, train.py:51-53).  With one process per tower
    the data-dependent ActNorm initialisation (train.py:221,229) would leave every rank with its own statistics, so rank `src`'s
    variable vector is broadcast after it (documented deviation -- the reference's init step lets the last tower win)."""
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1:
        torch.distributed.broadcast(flat, src=src, group=group)
    return flat


class Trainer:
    """One tower.  `group`: torch.distributed process group over which tower gradients are averaged (None = single tower)."""

    def __init__(self, model, group=None, clip_norm=1.0, beta1=0.9, beta2=0.999, epsilon=1e-8, split_terms=6, compute_dtype="float32",
                 exact_forward=None, overlap_allreduce=True):
        """model: an fp32 FloWaveNet (fp32 master variables, as utils.fp16_dtype_getter keeps them, utils.py:3-31).
        compute_dtype: 'float32' = fp32-accurate GEMMs (3-way bf16 split on the tensor cores; `split_terms` products per fp32 product:
        6 = every product down to 2^-24, 3 = ~2^-16 at half the tensor work); 'bfloat16' = bf16 operands and tape with fp32
        accumulation, reductions and gradients (the reference's mixed-precision training, utils.py:3-31, with bf16 for fp16).
        exact_forward (fp32 only; default: on for split_terms=6): forward GEMMs of the step on the CUDA-core engine -- the parity
        setting that keeps EVERY variable's gradient within 2e-4 of its own max-abs (see fwn_set_train_exact_forward).
        The static loss scale of the reference (hparams.scale = 64, train.py:62,75-77) exists for fp16 gradients; every mode here
        accumulates and stores gradients in fp32, so it is 1 and not a parameter."""
        if model._precision != _lib.FWN_FP32:
            raise ValueError("training keeps fp32 master variables: build the model with hparams.dtype='float32' and choose the "
                             "compute dtype with Trainer(compute_dtype=...)")
        if compute_dtype not in ("float32", "bfloat16"):
            raise ValueError("unsupported compute_dtype %r (float32 or bfloat16)" % (compute_dtype,))
        self.model, self.group = model, group
        self.clip_norm, self.beta1, self.beta2, self.epsilon = clip_norm, beta1, beta2, epsilon
        self.compute_dtype = compute_dtype
        self.global_step = 0
        L = _lib.lib()
        model._sync_params()
        with torch.cuda.device(model._device):
            _lib.check(L.fwn_train_enable(model._h, _lib.stream_ptr()))
        _lib.check(L.fwn_set_split_terms(model._h, 6, int(split_terms)))
        self.exact_forward = bool(split_terms == 6 and compute_dtype == "float32") if exact_forward is None else bool(exact_forward)
        _lib.check(L.fwn_set_train_exact_forward(model._h, int(self.exact_forward)))
        _lib.check(L.fwn_set_train_compute(model._h, _lib.FWN_MIXED_BF16 if compute_dtype == "bfloat16" else _lib.FWN_FP32))
        self._n = L.fwn_grad_floats(model._h)
        self._np = L.fwn_param_floats(model._h)
        self.grads = torch.zeros(self._n, dtype=torch.float32, device=model._device)
        self._ws = None
        self._out = torch.zeros(3, dtype=torch.float32, device=model._device)
        # gradient buckets (production order) for the overlapped tower average
        import ctypes as C
        off, cnt = C.c_int64(), C.c_int64()
        self.buckets = []
        for k in range(L.fwn_grad_bucket_count(model._h)):
            _lib.check(L.fwn_grad_bucket_range(model._h, k, C.byref(off), C.byref(cnt)))
            self.buckets.append((off.value, cnt.value))
        self.overlap_allreduce = bool(overlap_allreduce)
        self._comm = None
        self._pending = []

    def param_floats(self):
        return self._np

    def _towers(self):
        d = torch.distributed
        return d.get_world_size(self.group) if (d.is_available() and d.is_initialized()) else 1

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
        m._sync_params()   # variables loaded since the last step (FloWaveNet.load_variables) reach the handle; Adam slots are kept
        B, T = x.shape[0], x.shape[1]
        ws = self._workspace(B, T)
        with torch.cuda.device(m._device):
            _lib.check(_lib.lib().fwn_loss_and_grads(m._h, _lib.ptr(x), _lib.ptr(c), _lib.ptr(g), B, T, _lib.ptr(self._out[0:]),
                                                    _lib.ptr(self._out[1:]), _lib.ptr(self.grads), self._n, _lib.ptr(ws), ws.numel(),
                                                    _lib.stream_ptr()))
            if self.overlap_allreduce and self._towers() > 1:
                # the whole pass is enqueued; each bucket's all-reduce goes on the communication stream behind the event that marks
                # the bucket final, i.e. it runs while the backward pass of the blocks before it is still executing
                if self._comm is None:
                    self._comm = torch.cuda.Stream(device=m._device)
                L, h, comm = _lib.lib(), m._h, self._comm
                waiters = [(lambda k=k: _lib.check(L.fwn_grad_bucket_wait(h, k, ctypes.c_void_p(comm.cuda_stream)))) for k in range(len(self.buckets))]
                self._pending = average_flat_gradients_bucketed(self.grads, self.buckets, self.group, waiters, comm)
        return self._out[0], self._out[1]

    def average_gradients(self):
        """utils.py:34-60 across towers = processes.  With overlap_allreduce the reductions were enqueued by loss_and_grads; here the
        compute stream only waits for them."""
        if self._pending:
            for w in self._pending:
                w.wait()   # the current stream waits for the collective; the host does not block
            self._pending = []
            return
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

    def save_checkpoint(self, prefix, scope="vocoder/FloWaveNet"):
        """tf.train.Saver.save (train.py:190,252) in TensorFlow's own file format (checkpoint.py, no TensorFlow needed): the variables
        under `scope`, their Adam slots ('<name>/Adam', '<name>/Adam_1'), beta powers and global_step.  The reference's
        synthesize.py / train.py restore (and tf_flowavenet_b200.synthesize --saved_dir) read it back."""
        from . import checkpoint
        m, out = self.model, {}
        state = self.state_dict()
        flat = {k: state[k].cpu().numpy() for k in ("variables", "adam_m", "adam_v")}
        L = _lib.lib()
        for i, (k, shp) in enumerate(m.variable_shapes().items()):
            off = L.fwn_param_offset(m._h, i)
            n = 1
            for s in shp:
                n *= s
            name = scope + "/" + k
            out[name] = flat["variables"][off:off + n].reshape(shp)
            out[name + "/Adam"] = flat["adam_m"][off:off + n].reshape(shp)
            out[name + "/Adam_1"] = flat["adam_v"][off:off + n].reshape(shp)
        import numpy as np
        out["global_step"] = np.array(self.global_step, dtype=np.int64)
        out["beta1_power"] = np.array(self.beta1 ** self.global_step, dtype=np.float32)
        out["beta2_power"] = np.array(self.beta2 ** self.global_step, dtype=np.float32)
        return checkpoint.write_checkpoint(prefix, out)

    def restore_checkpoint(self, prefix, scope="vocoder/FloWaveNet"):
        """Saver.restore (train.py:211-219): variables, Adam slots and the step counter from a TensorBundle checkpoint -- one written by
        save_checkpoint or by the reference's own training run."""
        import numpy as np
        from . import checkpoint
        m, L = self.model, _lib.lib()
        ck = checkpoint.load_checkpoint(prefix)
        flat = {k: np.zeros(self._np, dtype=np.float32) for k in ("variables", "adam_m", "adam_v")}
        for i, (k, shp) in enumerate(m.variable_shapes().items()):
            off = L.fwn_param_offset(m._h, i)
            name = scope + "/" + k
            if name not in ck:
                raise KeyError("checkpoint %s has no variable '%s'" % (prefix, name))
            for key, suffix in (("variables", ""), ("adam_m", "/Adam"), ("adam_v", "/Adam_1")):
                if name + suffix in ck:     # the reference creates no Adam slot for variables without a gradient (speaker embeddings)
                    a = np.asarray(ck[name + suffix], dtype=np.float32).reshape(-1)
                    flat[key][off:off + a.size] = a
        state = {k: torch.from_numpy(v) for k, v in flat.items()}
        state["global_step"] = int(ck["global_step"]) if "global_step" in ck else 0
        self.load_state_dict(state)

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
        """{variable name -> fresh copy of the current value} (the handle owns the live variables during training;
        FloWaveNet.variables() returns the same values through the model's store)."""
        m, L, out = self.model, _lib.lib(), {}
        with torch.cuda.device(m._device):
            for k, shp in m.variable_shapes().items():
                t = torch.empty(shp, dtype=torch.float32, device=m._device)
                _lib.check(L.fwn_get_param(m._h, k.encode(), _lib.ptr(t), t.numel(), _lib.stream_ptr()))
                out[k] = t
        return out
