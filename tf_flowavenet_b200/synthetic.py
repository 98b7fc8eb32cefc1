"""Seeded synthetic weights and inputs for benchmarks and examples (no dataset / checkpoint is available offline).

Mirrors the reference initialisers (he-uniform kernels and biases, modules.py:21-22,79-97) but with wn/g ~ U(0.5,1.5)
and a small NON-zero ZeroConv1d so the couplings do real work (the reference zero-inits them, modules.py:46-49,
which would make every coupling the identity).  SURVEY Appendix A.7."""
import math

import numpy as np


def synthetic_params(shapes, seed=0):
    """shapes: {name -> shape} as returned by FloWaveNet.variable_shapes()."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shp in shapes.items():
        if name.endswith("/kernel"):
            if "ZeroConv1d" in name:
                a = rng.uniform(-0.02, 0.02, shp)
            else:
                fan_in = int(np.prod(shp[:-1])) if len(shp) == 3 else shp[0] * shp[1]
                lim = math.sqrt(6.0 / fan_in)
                a = rng.uniform(-lim, lim, shp)
        elif name.endswith("/wn/g"):
            a = rng.uniform(0.5, 1.5, shp)
        elif name.endswith("/bias"):
            if "ZeroConv1d" in name:
                a = rng.uniform(-0.02, 0.02, shp)
            elif name.startswith("conv2d_transpose"):
                a = rng.uniform(-0.05, 0.05, shp)
            else:
                a = rng.uniform(-1, 1, shp) * math.sqrt(6.0 / shp[0]) * 0.1
        elif name.endswith("/scale"):
            a = rng.uniform(-0.1, 0.1, shp)
        elif name.endswith("/ActNorm/b"):
            a = rng.uniform(-0.1, 0.1, shp)
        elif name.endswith("/ActNorm/logs"):
            a = rng.uniform(-0.05, 0.05, shp)
        elif name == "speaker_embeddings":
            a = rng.standard_normal(shp) * 0.1
        else:
            raise KeyError(name)
        out[name] = a.astype(np.float32)
    return out


def synthetic_inputs(hop, num_mels, B, n_frames, seed, kind="z", temp=0.7):
    """mel ~ U[0,1) (range of preprocessing.py:68-69); z ~ N(0,1)*temp (synthesize.py:14); x: chirp + noise in (-1,1)."""
    rng = np.random.default_rng(seed)
    T = n_frames * hop
    c = rng.uniform(0.0, 1.0, (B, n_frames, num_mels)).astype(np.float32)
    if kind == "z":
        a = (rng.standard_normal((B, T, 1)) * temp).astype(np.float32)
    else:
        t = np.arange(T, dtype=np.float64)[None, :, None] / T
        ph = rng.uniform(0, 2 * np.pi, (B, 1, 1))
        a = 0.5 * np.sin(2 * np.pi * (200.0 * t + 800.0 * t * t) + ph) + 0.1 * rng.standard_normal((B, T, 1))
        a = np.clip(a, -0.999, 0.999).astype(np.float32)
    return a, c
