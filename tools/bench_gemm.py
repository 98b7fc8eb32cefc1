"""Micro-benchmark of the tcgen05 implicit-GEMM kernel through fwn_conv1d_bf16 (EPI_PLAIN): separates the MMA/operand
path ceiling (long K, negligible epilogue) from epilogue / TMEM hand-off effects (short K)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tf_flowavenet_b200 import _lib

L = _lib.lib()


def run(B, T, Cin, Cout, k=1, d=1, iters=20):
    Cin16, Npad = (Cin + 15) // 16 * 16, (Cout + 15) // 16 * 16
    Kpad = (k * Cin16 + 63) // 64 * 64
    x = torch.randn(B, T, Cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(Npad, Kpad, device="cuda") / (k * Cin) ** 0.5).to(torch.bfloat16)
    bias = torch.zeros(Cout, device="cuda")
    y = torch.empty(B, T, Cout, device="cuda", dtype=torch.bfloat16)
    f = lambda: _lib.check(L.fwn_conv1d_bf16(_lib.ptr(x), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(y), B, T, Cin, Cout, k, d, 0, 0, None))
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * B * T * k * Cin * Cout
    print("B=%d T=%d Cin=%d Cout=%d k=%d: %.3f ms  %.1f TFLOP/s  (in+out %.1f GB/s)" % (B, T, Cin, Cout, k, ms, flops / ms / 1e9,
          (x.numel() + y.numel()) * 2 / ms / 1e6))


if __name__ == "__main__":
    run(8, 32768, 4096, 256)      # long K: MMA / operand path ceiling
    run(8, 32768, 2048, 256)
    run(8, 32768, 1024, 256)
    run(32, 40032, 256, 256, k=3)  # gate-like K=768 without the gate epilogue, N=256
    run(32, 40032, 256, 256, k=1)  # K=256
    run(8, 32768, 4096, 128)      # BN=128
