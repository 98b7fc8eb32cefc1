"""Training step in the bf16 compute mode (BASELINE config 5: "bf16"; the reference's mixed-precision training, utils.py:3-31 +
train.py:53-81, with bf16 for fp16 and loss scale 1) through the C ABI: fwn_set_train_compute(FWN_MIXED_BF16) + fwn_loss_and_grads.

Referees: the float64 training oracle on the small configurations; the library's own fp32 parity mode (itself held to the oracle in
test_gpu_train.py) at the full hparams.py depth.  Stated bounds.  bf16 operands carry 2^-9 relative rounding on every GEMM input, and a pre-activation that lands on the other side of
zero flips a ReLU mask outright, so single entries of a gradient that sums over a few dozen rows can be off by 10 % of the variable's
largest entry (measured 1.4e-1 on the 16-row fixture, where the whole vector is within 1.7e-3); errors are therefore bounded in L2:
  every variable with >= 512 elements:  ||g - g_ref|| <= 1.5e-1 * ||g_ref||   (measured up to 1e-1)   (floored at 1e-3 of the model's RMS gradient)
  whole vector:                         relative L2 error <= 4e-2   (measured 2e-3 ... 2e-2)
The upsampler's 98 + 98 parameters (sums of the conditioning gradient over every sample and mel bin, cancellation-dominated) are
reported, not bounded: measured up to 4e-1 relative -- train them in the fp32 mode if they matter.
Each test prints what it measured, including the worst single entry."""
import numpy as np
import pytest
import torch

from oracle import flowavenet_oracle as O
from oracle import flowavenet_train_oracle as TO
from tests._golden import load
from tests.test_gpu_model import make_model

pytestmark = pytest.mark.gpu


def grad_errors(got, ref, min_numel=512):
    """-> (worst per-variable relative L2 error over the variables with >= min_numel elements, its name), worst single-entry error
    relative to its variable's max, relative L2 of the whole vector.  Small variables are floored at 1e-3 of the model's scale.  The
    five worst variables of any size are printed: the 98-parameter upsampler kernels / scalars are sums of the conditioning gradient
    over every sample and mel bin with heavy cancellation and carry the largest relative errors."""
    gmax = max(float(r.abs().max()) for r in ref.values())
    n_all = sum(r.numel() for r in ref.values())
    rms_all = (sum(float((r.double() ** 2).sum()) for r in ref.values()) / n_all) ** 0.5
    worst, worst_max, rows = ("", 0.0), 0.0, []
    num = den = 0.0
    for k, r in ref.items():
        g, r = got[k].double().cpu(), r.double().cpu()
        e2, r2 = float(((g - r) ** 2).sum()), float((r ** 2).sum())
        err = (e2 / r.numel()) ** 0.5 / max((r2 / r.numel()) ** 0.5, 1e-3 * rms_all)
        rows.append((err, k, r.numel(), (r2 / r.numel()) ** 0.5 / rms_all))
        if err > worst[1] and r.numel() >= min_numel:
            worst = (k, err)
        worst_max = max(worst_max, float((g - r).abs().max()) / max(float(r.abs().max()), 1e-3 * gmax))
        num += e2
        den += r2
    for err, k, n, rel in sorted(rows, reverse=True)[:5]:
        print("    %-70s numel %-8d rms/model-rms %.2e  rel-L2 err %.2e" % (k, n, rel, err))
    return worst, worst_max, (num / den) ** 0.5


@pytest.mark.parametrize("B,T,K,N,shift", [(2, 300, 256, 256, 0), (3, 130, 80, 512, -3), (1, 64, 8, 16, 1), (2, 1000, 264, 128, 9),
                                           (4, 25, 10240, 512, 0), (1, 2000, 768, 512, -1)])
def test_wgrad_bf16_vs_torch(B, T, K, N, shift):
    """One launch of the bf16 weight-gradient kernel (TMA -> MN-major tcgen05.mma -> TMEM -> fp32 atomics) against an fp32 einsum
    on the same bf16 inputs, including the time shift of a dilated tap (rows outside [0, T) read as zero) and the bias gradient."""
    from tf_flowavenet_b200 import _lib
    rng = np.random.default_rng(B + T + K + N)
    a = torch.from_numpy(rng.standard_normal((B, T, K))).float().to(torch.bfloat16)
    dy = torch.from_numpy(rng.standard_normal((B, T, N))).float().to(torch.bfloat16)
    ad, dyd = a.cuda().contiguous(), dy.cuda().contiguous()
    dw = torch.zeros(K, N, device="cuda")
    db = torch.zeros(N, device="cuda")
    _lib.check(_lib.lib().fwn_wgrad_bf16(_lib.ptr(ad), _lib.ptr(dyd), _lib.ptr(dw), _lib.ptr(db), B, T, K, N, shift, None))
    torch.cuda.synchronize()
    af = torch.zeros(B, T, K)
    lo, hi = max(0, -shift), min(T, T - shift)          # rows t with 0 <= t + shift < T
    af[:, lo:hi] = a.float()[:, lo + shift:hi + shift]
    want = torch.einsum("btk,btn->kn", af.double(), dy.double())
    err = float((dw.cpu().double() - want).abs().max() / want.abs().max())
    berr = float((db.cpu().double() - dy.double().sum((0, 1))).abs().max() / dy.double().sum((0, 1)).abs().max())
    print("wgrad bf16 B=%d T=%d K=%d N=%d shift=%d: dW rel-to-max err %.2e, dbias %.2e" % (B, T, K, N, shift, err, berr))
    assert err < 1e-4 and berr < 1e-4


@pytest.mark.parametrize("case", ["g1_b2f2l2"])
def test_bf16_gradients_small_vs_oracle(case):
    import tf_flowavenet_b200.train as T
    hp, params, fx = load(case)
    tr = T.Trainer(make_model(hp, params), compute_dtype="bfloat16")
    x, c = torch.from_numpy(fx["x"]).float().cuda(), torch.from_numpy(fx["c"]).float().cuda()
    log_p, logdet = tr.loss_and_grads(x, c)
    assert abs(float(log_p) - float(fx["log_p"])) < 1e-2 and abs(float(logdet) - float(fx["logdet"])) < 1e-2
    _, _, _, ref = TO.loss_and_grads(params, hp, torch.from_numpy(fx["x"]), torch.from_numpy(fx["c"]))
    worst, wmax, l2 = grad_errors(tr.gradients(), ref)
    print("bf16 gradients %s: worst per-variable L2 %.2e (%s), worst single entry %.2e of its variable's max, whole-vector L2 %.2e" %
          (case, worst[1], worst[0], wmax, l2))
    assert worst[1] < 1.5e-1 and l2 < 4e-2


@pytest.mark.parametrize("B,n_frames", [(1, 160), (3, 100)])
def test_bf16_gradients_multi_tile_vs_oracle(B, n_frames):
    """Several (partial) 128-row tiles and 64-step wgrad chunks, several utterances, ActNorm from the data."""
    import tf_flowavenet_b200.train as T
    hp = O.HP(n_block=2, n_flow=2, n_layer=2, num_mels=8, upsample_scales=(2, 2))
    params = O.synthetic_params(hp, seed=21, dtype=torch.float64)
    x, c = O.synthetic_inputs(hp, B, n_frames, 22, "x")
    params = O.ddi_init(params, hp, x, c, torch.float64)
    loss, _, _, ref = TO.loss_and_grads(params, hp, x, c)
    tr = T.Trainer(make_model(hp, params), compute_dtype="bfloat16")
    log_p, logdet = tr.loss_and_grads(x.float().cuda(), c.float().cuda())
    assert abs(float(-(log_p + logdet)) - loss) < 1e-2 * max(1.0, abs(loss))
    worst, wmax, l2 = grad_errors(tr.gradients(), ref)
    print("bf16 gradients B=%d frames=%d: worst per-variable L2 %.2e (%s), worst single entry %.2e, whole-vector L2 %.2e" %
          (B, n_frames, worst[1], worst[0], wmax, l2))
    assert worst[1] < 1.5e-1 and l2 < 4e-2


def test_bf16_gradients_full_depth_c5_shape_vs_fp32_mode():
    """hparams.py (8x6x2, gin 16) at the per-GPU C5 shape (8 x 6400): bf16 step against the library's fp32 parity step."""
    import tf_flowavenet_b200 as P
    import tf_flowavenet_b200.train as T
    from tf_flowavenet_b200.synthetic import synthetic_inputs, synthetic_params
    net = P.FloWaveNet(P.HParams(**{**P.hparams.values(), "dtype": "float32", "gin_channels": 16, "n_speakers": 7}), variables=P.VariableStore())
    net.load_variables(synthetic_params(net.variable_shapes(), seed=5))
    x, c = synthetic_inputs(256, 80, 8, 25, 6, "x")
    x, c = torch.from_numpy(x).cuda(), torch.from_numpy(c).cuda()
    g = torch.arange(8, dtype=torch.int32, device="cuda") % 7
    net.initialize_actnorm(x, c, g)
    tr = T.Trainer(net, split_terms=6, exact_forward=False)
    lp32, ld32 = (float(v) for v in tr.loss_and_grads(x, c, g))
    ref = {k: v.clone() for k, v in tr.gradients().items()}
    tr16 = T.Trainer(net, compute_dtype="bfloat16")
    lp16, ld16 = (float(v) for v in tr16.loss_and_grads(x, c, g))
    worst, wmax, l2 = grad_errors(tr16.gradients(), ref)
    print("bf16 vs fp32 step, hparams 8x6400: log_p %.5f vs %.5f, logdet %.5f vs %.5f; worst per-variable L2 %.2e (%s), worst single entry "
          "%.2e, whole-vector L2 %.2e; %d launches" % (lp16, lp32, ld16, ld32, worst[1], worst[0], wmax, l2, net.last_launches()))
    assert abs(lp16 - lp32) < 2e-2 and abs(ld16 - ld32) < 2e-2
    assert worst[1] < 1.5e-1 and l2 < 4e-2


def test_bf16_training_trajectory():
    """Five optimizer steps in bf16 compute follow the fp32 trajectory (same variables, same batch) and reduce the loss."""
    import tf_flowavenet_b200.train as T
    hp, params, fx = load("g1_b2f2l2")
    x, c = torch.from_numpy(fx["x"]).float().cuda(), torch.from_numpy(fx["c"]).float().cuda()
    tr32, tr16 = T.Trainer(make_model(hp, params)), T.Trainer(make_model(hp, params), compute_dtype="bfloat16")
    l32 = [float(tr32.train_step(x, c)["loss"]) for _ in range(5)]
    l16 = [float(tr16.train_step(x, c)["loss"]) for _ in range(5)]
    print("loss trajectory fp32 %s | bf16 %s" % (["%.5f" % v for v in l32], ["%.5f" % v for v in l16]))
    assert l16[-1] < l16[0]
    assert max(abs(a - b) for a, b in zip(l32, l16)) < 2e-2


def test_fp32_passes_after_bf16_steps_use_fresh_operands():
    """The bf16 optimizer step refreshes only plane 0 of the split engine's three operand planes (it reads nothing else); the fp32
    passes of the same handle (model.py:317-396 on the trained variables) must see all three refreshed.  After three bf16 steps the fp32
    forward equals the float64 oracle on the CURRENT variables (1e-4), and an explicit full re-pack changes nothing."""
    import tf_flowavenet_b200.train as T
    from tf_flowavenet_b200 import _lib
    hp, params, fx = load("g1_b2f2l2")
    x, c = torch.from_numpy(fx["x"]).float().cuda(), torch.from_numpy(fx["c"]).float().cuda()
    net = make_model(hp, params)
    tr = T.Trainer(net, compute_dtype="bfloat16")
    for _ in range(3):
        tr.train_step(x, c)
    lp, ld, z = net.forward(x, c, return_z=True)
    cur = {k: v.detach().cpu() for k, v in net.variables().items()}
    wlp, wld, wz = O.forward(cur, hp, x.cpu(), c.cpu(), torch.float64)
    err = float((z.cpu().double() - wz).abs().max() / wz.abs().max())
    print("fp32 forward after 3 bf16 steps: z rel-to-max %.2e, log_p %.6f vs %.6f" % (err, float(lp), float(wlp)))
    assert err < 1e-4 and abs(float(lp) - float(wlp)) < 1e-4 * max(1.0, abs(float(wlp)))
    _lib.check(_lib.lib().fwn_repack(net._h, _lib.stream_ptr()))
    lp2, ld2, z2 = net.forward(x, c, return_z=True)
    assert torch.equal(z, z2)
    # and the fp32 training mode on the same handle after bf16 steps: gradients of the current variables against the oracle
    tr.train_step(x, c)
