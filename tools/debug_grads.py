#!/usr/bin/env python
"""Per-variable gradient error listing (CUDA training step vs the float64 oracle) for a small model; debugging aid.
   python tools/debug_grads.py B n_frames [n_block]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import flowavenet_oracle as O  # noqa: E402
from oracle import flowavenet_train_oracle as TO  # noqa: E402
from tests.test_gpu_model import make_model  # noqa: E402
import tf_flowavenet_b200.train as T  # noqa: E402

B, nf = int(sys.argv[1]), int(sys.argv[2])
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 2
hp = O.HP(n_block=nb, n_flow=2, n_layer=2, num_mels=8, upsample_scales=(2, 2))
params = O.synthetic_params(hp, seed=21, dtype=torch.float64)
x, c = O.synthetic_inputs(hp, B, nf, 22, "x")
params = O.ddi_init(params, hp, x, c, torch.float64)
tr = T.Trainer(make_model(hp, params), split_terms=int(os.environ.get('TERMS', '3')))
log_p, logdet = tr.loss_and_grads(x.float().cuda(), c.float().cuda())
loss, _, _, ref = TO.loss_and_grads(params, hp, x, c)
print("loss", float(-(log_p + logdet)), loss)
got = tr.gradients()
gmax = max(float(r.abs().max()) for r in ref.values())
for k, r in ref.items():
    g = got[k].double().cpu()
    err = float((g - r).abs().max()) / max(float(r.abs().max()), 1e-6 * gmax)
    if err > 1e-4:
        print("%-75s err %.3e  |ref|max %.3e" % (k, err, float(r.abs().max())))
