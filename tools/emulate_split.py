#!/usr/bin/env python
"""CPU emulation of the bf16x3 split GEMM (csrc/gemm_tc3.cu) on a gate-GEMM-shaped problem: relative error against float64 of
  - a plain fp32 GEMM (round-to-nearest FMA chain),
  - the 3- and 6-term split with exact fp32 accumulation of the bf16 products (what the split alone costs),
  - the same with an accumulator that TRUNCATES (round toward zero) when a 16-product MMA partial sum is added to the running sum
    (a model of the tensor-core accumulator; reproduces the ~5e-6 measured on B200 for 6 terms, K = 848).
DESIGN.md section 4 quotes these numbers.   python tools/emulate_split.py [M K N]"""
import sys

import numpy as np
import torch


def bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def split3(x):
    a1 = bf16(x)
    a2 = bf16(x - a1)
    a3 = bf16(x - a1 - a2)
    return a1, a2, a3


def trunc_add(acc, part):
    """fp32(acc + part) rounded toward zero instead of to nearest."""
    s = acc.double() + part.double()
    r = s.float()                                   # round to nearest
    over = (r.double().abs() > s.abs())              # rounded away from zero -> step one ulp back towards zero
    r = torch.where(over, torch.nextafter(r, torch.zeros_like(r)), r)
    return r


def gemm_terms(A, W, terms, truncate):
    a, w = split3(A), split3(W)
    pairs = [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)][:terms]
    M, K = A.shape
    acc = torch.zeros(M, W.shape[1], dtype=torch.float32)
    for k0 in range(0, K, 16):                       # one MMA = 16 products per output
        for pa, pw in pairs:
            part = (a[pa][:, k0:k0 + 16].double() @ w[pw][k0:k0 + 16].double()).float()   # products exact, partial sum ~exact
            acc = trunc_add(acc, part) if truncate else (acc.double() + part.double()).float()
    return acc


def main():
    M, K, N = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (256, 848, 128)
    g = torch.Generator().manual_seed(0)
    A = torch.randn(M, K, generator=g) * 1.5
    W = (torch.rand(K, N, generator=g) - 0.5) * 0.2
    ref = A.double() @ W.double()
    scale = float(ref.abs().max())

    def err(x):
        return float((x.double() - ref).abs().max()) / scale

    print("shape M=%d K=%d N=%d; errors are max-abs / max|result|" % (M, K, N))
    print("fp32 FMA chain (torch CPU)            %.2e" % err(A @ W))
    for terms in (3, 6):
        print("%d-term split, exact accumulation      %.2e" % (terms, err(gemm_terms(A, W, terms, False))))
        print("%d-term split, truncating accumulator  %.2e" % (terms, err(gemm_terms(A, W, terms, True))))


if __name__ == "__main__":
    main()
