"""CPU restatement of the reference's TRAINING STEP -- test infrastructure only (see flowavenet_oracle.py's header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

What it restates (reference file:line):
  * loss = -(log_p + logdet), gradients of the loss w.r.t. every trainable variable      train.py:59-66
    (torch autograd over the oracle's forward graph, which is pinned to the reference's own Python by the golden fixtures)
  * average of the tower gradients                                                         utils.py:34-60
  * tf.clip_by_global_norm(grads, 1)                                                       train.py:27-32
  * learning-rate schedule 1e-3, /2, /4, /6 at 200k/400k/600k steps                        train.py:15-24
  * tf.train.AdamOptimizer(lr) defaults beta1=.9 beta2=.999 eps=1e-8                       train.py:22, [TF] adam.py
Parity status: PINNED.  tests/golden/*_grads.npz hold tf.gradients(loss, tf.trainable_variables()) of the reference's own
model.py (imported unmodified, train.py:59-63 replayed literally) run on the eager TF shim (tests/golden/make_golden_grads.py);
tests/test_train_oracle.py checks this oracle against them (norm and probed entries of every variable's gradient, global norm)
and against central finite differences.  TensorFlow 1.12 itself is not installable here, so "reference" means the reference's
Python graph code over the shim's restatement of the TF ops (same status as the forward fixtures).
"""
from typing import Dict, Tuple

import torch

from . import flowavenet_oracle as fo

Params = Dict[str, torch.Tensor]


def loss_and_grads(p: Params, hp, x: torch.Tensor, c: torch.Tensor, dtype=torch.float64) -> Tuple[float, float, float, Params]:
    """One tower of build_model (train.py:56-66).  Variables the loss does not depend on get a zero gradient
    (tf.gradients returns None for them and train.py:75 drops them; Adam then leaves them untouched, same as zero)."""
    leaf = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in p.items()}
    log_p, logdet, _ = fo.forward(leaf, hp, x, c, dtype=dtype)
    loss = -(log_p + logdet)
    names = list(leaf)
    gs = torch.autograd.grad(loss, [leaf[k] for k in names], allow_unused=True)
    grads = {k: (torch.zeros_like(leaf[k]) if g is None else g.detach()) for k, g in zip(names, gs)}
    return float(loss.detach()), float(log_p.detach()), float(logdet.detach()), grads


def average_gradients(tower_grads):
    """utils.py:34-60: mean over towers, variable by variable."""
    return {k: torch.stack([tg[k] for tg in tower_grads], 0).mean(0) for k in tower_grads[0]}


def global_norm(grads: Params) -> float:
    return float(torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())))


def clip_by_global_norm(grads: Params, clip: float = 1.0) -> Tuple[Params, float]:
    """train.py:27-32 / tf.clip_by_global_norm: g * clip / max(norm, clip)."""
    n = global_norm(grads)
    s = clip / max(n, clip)
    return {k: g * s for k, g in grads.items()}, n


def learning_rate(global_step: int) -> float:
    """train.py:15-20 (three nested tf.cond on the step)."""
    lr = 0.001
    if not global_step < 200000:
        lr = 0.001 / 2
    if not global_step < 400000:
        lr = 0.001 / 4
    if not global_step < 600000:
        lr = 0.001 / 6
    return lr


def adam_step(p: Params, grads: Params, m: Params, v: Params, lr: float, t: int, b1=0.9, b2=0.999, eps=1e-8):
    """[TF] training/adam.py: lr_t = lr sqrt(1-b2^t)/(1-b1^t); m,v moving averages; p -= lr_t m / (sqrt(v) + eps).  t counts from 1."""
    lr_t = lr * (1 - b2 ** t) ** 0.5 / (1 - b1 ** t)
    out = {}
    for k in p:
        g = grads[k].to(p[k].dtype)
        m[k] = b1 * m[k] + (1 - b1) * g
        v[k] = b2 * v[k] + (1 - b2) * g * g
        out[k] = p[k] - lr_t * m[k] / (v[k].sqrt() + eps)
    return out


def train_step(p: Params, hp, towers, m: Params, v: Params, global_step: int, dtype=torch.float64, clip: float = 1.0):
    """build_model + train_op (train.py:35-81) for `towers` = [(x, c), ...]; returns (new params, info)."""
    res = [loss_and_grads(p, hp, x, c, dtype) for x, c in towers]
    grads = average_gradients([r[3] for r in res])
    clipped, norm = clip_by_global_norm(grads, clip)
    lr = learning_rate(global_step)
    pp = {k: t.to(dtype) for k, t in p.items()}
    new = adam_step(pp, clipped, m, v, lr, global_step + 1)
    return new, {"loss": res[0][0], "log_p": res[0][1], "logdet": res[0][2], "grad_global_norm": norm, "lr": lr}
