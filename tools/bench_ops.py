"""GB/s of the bandwidth-bound per-op C-ABI kernels (SURVEY 8d algorithmic bytes) against the measured HBM copy peak.
Tensors are > 126 MB (L2) so every repetition streams from HBM.  Prints one JSON line per op."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tf_flowavenet_b200 import _lib

L = _lib.lib()
PEAK = 6550.7
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    PEAK = json.load(open(p))["hbm_gbs"]


from bench import ClockSampler  # nvidia-smi clocks / throttle reasons sampled DURING each op's timed region

LAST_CLOCKS = None


def timeit(fn, iters=200):
    """ms per call; >= 200 iterations so the nvidia-smi sampler (100 ms period) sees the op under load."""
    global LAST_CLOCKS
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    cs = ClockSampler(0)
    cs.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    if ms * iters < 250:   # keep the op running long enough for a few clock samples
        extra = int(250 / max(ms, 1e-3))
        for _ in range(extra):
            fn()
        torch.cuda.synchronize()
    LAST_CLOCKS = cs.stop()
    return ms


def report(name, ms, nbytes, shape):
    gbs = nbytes / ms / 1e6
    print(json.dumps({"op": name, "shape": shape, "ms": round(ms, 4), "algorithmic_GB": round(nbytes / 1e9, 3), "GBps": round(gbs, 1),
                      "frac_of_measured_hbm_peak": round(gbs / PEAK, 3), "clocks": LAST_CLOCKS}), flush=True)


def main():
    dev = "cuda"
    B, T = 32, 1 << 21  # 67M samples = 268 MB fp32
    for C in (2, 8, 256):
        x = torch.randn(B, T // C, C, device=dev)
        y = torch.empty_like(x)
        n = x.numel()
        b, logs = torch.randn(C, device=dev) * 0.1, torch.randn(C, device=dev) * 0.05
        ld = torch.empty((), device=dev)
        rows = B * (T // C)
        report("squeeze C=%d" % C, timeit(lambda: _lib.check(L.fwn_squeeze(_lib.ptr(x), _lib.ptr(y), B, T // C, C, None))), 8 * n, [B, T // C, C])
        report("change_order C=%d" % C, timeit(lambda: _lib.check(L.fwn_change_order(_lib.ptr(x), _lib.ptr(y), rows, C, None))), 8 * n, [rows, C])
        report("actnorm_fwd C=%d" % C, timeit(lambda: _lib.check(L.fwn_actnorm_fwd(_lib.ptr(x), _lib.ptr(b), _lib.ptr(logs), _lib.ptr(y), _lib.ptr(ld), rows, C, None))), 8 * n, [rows, C])
        report("actnorm_rev C=%d" % C, timeit(lambda: _lib.check(L.fwn_actnorm_rev(_lib.ptr(x), _lib.ptr(b), _lib.ptr(logs), _lib.ptr(y), rows, C, None))), 8 * n, [rows, C])
        net = torch.randn(B, T // C, C, device=dev) * 0.1
        # affine: read x (all), net (all), write y (all) = 12 n bytes moved; SURVEY counts 2*E*s for the transformed half only -> report moved bytes
        report("affine_fwd C=%d" % C, timeit(lambda: _lib.check(L.fwn_affine_fwd(_lib.ptr(x), _lib.ptr(net), _lib.ptr(y), _lib.ptr(ld), rows, C, 1, None))), 12 * n, [rows, C])
        report("affine_rev C=%d" % C, timeit(lambda: _lib.check(L.fwn_affine_rev(_lib.ptr(x), _lib.ptr(net), _lib.ptr(y), rows, C, 1, None))), 12 * n, [rows, C])
        del x, y, net
    # squeeze of the conditioning tensor (80 channels): what the reference does per block
    c = torch.randn(8, 1 << 19, 80, device=dev)
    yc = torch.empty_like(c)
    report("squeeze C=80 (cond)", timeit(lambda: _lib.check(L.fwn_squeeze(_lib.ptr(c), _lib.ptr(yc), 8, 1 << 19, 80, None))), 8 * c.numel(), list(c.shape))
    del c, yc
    # upsampler stage (s=16): read B*Tm*80*4, write 16x that
    Bm, Tm = 32, 8192
    cin = torch.rand(Bm, Tm, 80, device=dev)
    k, g, bias = torch.randn(32, 3, 1, 1, device=dev), torch.ones(1, device=dev), torch.zeros(1, device=dev)
    cout = torch.empty(Bm, Tm * 16, 80, device=dev)
    report("upsample_stage s=16", timeit(lambda: _lib.check(L.fwn_upsample_stage(_lib.ptr(cin), _lib.ptr(k), _lib.ptr(g), _lib.ptr(bias), _lib.ptr(cout), Bm, Tm, 80, 16, None))),
           4 * (cin.numel() + cout.numel()), [Bm, Tm, 80])
    z = torch.randn(1 << 27, device=dev)
    out = torch.empty((), device=dev)
    report("log_p", timeit(lambda: _lib.check(L.fwn_log_p(_lib.ptr(z), _lib.ptr(out), z.numel(), None))), 4 * z.numel(), [z.numel()])


if __name__ == "__main__":
    main()
