#!/usr/bin/env python
"""Debugging aid: run the training forward+backward twice in one process -- all GEMMs on the split engine, then with some GEMM
families forced onto the CUDA-core engine (FWN_SIMT_FAMILIES) -- and diff the workspace (tape + scratch) region by region.
   python tools/debug_tape.py B n_frames mask"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import flowavenet_oracle as O  # noqa: E402
from tests.test_gpu_model import make_model  # noqa: E402
import tf_flowavenet_b200.train as T  # noqa: E402

B, nf, mask = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
hp = O.HP(n_block=2, n_flow=2, n_layer=2, num_mels=8, upsample_scales=(2, 2))
params = O.synthetic_params(hp, seed=21, dtype=torch.float64)
x, c = O.synthetic_inputs(hp, B, nf, 22, "x")
params = O.ddi_init(params, hp, x, c, torch.float64)
tr = T.Trainer(make_model(hp, params), split_terms=6)
xs, cs = x.float().cuda(), c.float().cuda()
os.environ["FWN_TRAIN_STREAMS"] = "1"


def run(m):
    os.environ["FWN_SIMT_FAMILIES"] = m
    tr._ws = None
    ws = tr._workspace(B, xs.shape[1])
    ws.zero_()
    tr.loss_and_grads(xs, cs)
    torch.cuda.synchronize()
    return tr._ws.clone().view(torch.float32), tr.grads.clone()


w0, g0 = run("0")
w1, g1 = run(mask)
# replicate train_plan's layout (csrc/train.cu)
BT, F, H, L = B * xs.shape[1], 256, 4, 2
M0 = BT // 2
off = 0
regions = []


def take(name, floats):
    global off
    regions.append((name, off // 4, floats))
    off = (off + floats * 4 + 255) & ~255


take("sums", 16); take("ddi", 2 * 4096 * 2); take("X", BT); take("dX", BT)
up = B * (xs.shape[1] // 2) * 8
take("up0", up); take("dup0", up)
for n in ("cA", "cB", "dcA", "dcB"):
    take(n, BT * H)
for s in range(2):
    take("set%d.dnet" % s, 2 * BT); take("set%d.da0" % s, 2 * BT); take("set%d.du" % s, M0 * F); take("set%d.ds" % s, M0 * F)
    for n in range(L):
        take("set%d.dfg%d" % (s, n), M0 * 2 * F); take("set%d.r%d" % (s, n), M0 * F)
for i in range(2):
    M = BT >> (i + 1)
    nq = 1 << i
    for j in range(2):
        p = "b%df%d." % (i, j)
        take(p + "xpre", BT); take(p + "a0", M * ((nq + 3) // 4 * 4))
        for n in range(L):
            take(p + "h%d" % n, M * F); take(p + "fg%d" % n, M * 2 * F); take(p + "o%d" % n, M * F)
        take(p + "s", M * F); take(p + "u", M * F); take(p + "net", M * ((2 * nq + 3) // 4 * 4))
assert off == w0.numel() * 4, (off, w0.numel() * 4)
for name, o, n in regions:
    a, b = w0[o:o + n].double(), w1[o:o + n].double()
    d = float((a - b).abs().max())
    sc = float(b.abs().max())
    flag = "  <<<<" if d > 1e-4 * max(sc, 1e-30) else ""
    print("%-14s max|diff| %.3e  max|val| %.3e  rel %.2e%s" % (name, d, sc, d / max(sc, 1e-30), flag))
print("grads rel diff", float((g0 - g1).abs().max() / g1.abs().max()))
