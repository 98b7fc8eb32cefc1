"""Generate tests/golden/*_grads.npz: gradients of the training loss computed by the REFERENCE'S OWN Python on the eager TF shim.

    python tests/golden/make_golden_grads.py        (build container only: needs /root/reference)

What runs: /root/reference/model.py (unmodified) builds the graph; then, literally as train.py:59-63,
    loss = -(log_p + logdet);  variables = tf.trainable_variables();  grads = tf.gradients(loss, variables)
with tf.gradients = torch autograd over the shim's eagerly recorded ops.  Stored per variable (fixtures stay small):
the L2 norm of its gradient and 6 entries at seeded positions; plus the global norm (train.py:29).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import tensorflow as tf  # noqa: E402  (the shim)
from oracle import flowavenet_oracle as O  # noqa: E402
from tests.golden.make_golden import CASES  # noqa: E402

GRAD_CASES = ["g1_b2f2l2", "g2_b3f2l1", "g4_causal", "g5_additive", "g6_l3"]
NPROBE = 6


def main():
    for name in GRAD_CASES:
        kw, B, nf, seed = CASES[name]
        hp = O.HP(**kw)
        params = O.synthetic_params(hp, seed=seed, dtype=torch.float64)
        x, c = O.synthetic_inputs(hp, B, nf, seed + 100, "x")
        tf.reset_default_graph()
        for m in ("model", "modules", "convolutional", "utils"):
            sys.modules.pop(m, None)
        import model as ref_model  # /root/reference/model.py
        tf.set_presets({"FloWaveNet/" + k: v.numpy() for k, v in params.items()})
        tf.set_trainable(True)
        hparams = tf.contrib.training.HParams(
            n_block=hp.n_block, n_flow=hp.n_flow, n_layer=hp.n_layer, num_mels=hp.num_mels, affine=hp.affine,
            causality=hp.causality, upsample_scales=list(hp.upsample_scales), gin_channels=-1, n_speakers=7, dtype=torch.float64)
        net = ref_model.FloWaveNet(hparams, init=False, scope="FloWaveNet")
        log_p, logdet = net.forward(tf.convert_to_tensor(x.double()), tf.convert_to_tensor(c.double()))
        loss = -(log_p + logdet)                                   # train.py:59
        named = {k[len("FloWaveNet/"):]: v for k, v in tf.global_variables_dict().items()}
        variables = tf.trainable_variables()                       # train.py:61
        assert len(variables) == len(named) == len(params), (len(variables), len(named), len(params))
        grads = tf.gradients(tf.scalar_mul(1.0, loss), variables)  # train.py:62-63 (hparams.scale = 1)
        tf.set_trainable(False)
        by_id = {id(v): g for v, g in zip(variables, grads)}
        rng = np.random.default_rng(seed + 300)
        out = {"loss": float(loss), "log_p": float(log_p), "logdet": float(logdet)}
        sq = 0.0
        for k in sorted(named):
            g = by_id[id(named[k])]
            g = torch.zeros_like(named[k]) if g is None else g
            g = g.detach().as_subclass(torch.Tensor).reshape(-1).double().numpy()
            idx = rng.choice(g.size, size=min(NPROBE, g.size), replace=False)
            out["norm::" + k] = np.float64(np.sqrt((g * g).sum()))
            out["idx::" + k] = idx.astype(np.int64)
            out["val::" + k] = g[idx]
            sq += float((g * g).sum())
        out["global_norm"] = np.float64(np.sqrt(sq))
        path = os.path.join(HERE, name + "_grads.npz")
        np.savez_compressed(path, **out)
        print(name, "loss", out["loss"], "global_norm", out["global_norm"], os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
