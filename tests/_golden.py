"""Helpers shared by CPU and GPU tests: load committed golden fixtures (tests/golden/*.npz)."""
import ast
import glob
import os

import numpy as np
import torch

from oracle import flowavenet_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")) if not p.endswith("_grads.npz"))


def load(name):
    fx = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    kw = ast.literal_eval(str(fx["hp"]))
    hp = O.HP(**kw)
    params = O.synthetic_params(hp, seed=int(fx["seed"]), dtype=torch.float64)
    chk = np.array([sum(float(v.sum()) for v in params.values()), sum(float((v * v).sum()) for v in params.values())])
    np.testing.assert_allclose(chk, fx["weight_checksum"], rtol=1e-12, err_msg="seeded weights drifted from the fixture")
    return hp, params, fx
