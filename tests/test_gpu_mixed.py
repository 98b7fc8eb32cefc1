"""Mixed-precision (bf16 operands, fp32 accumulate, tcgen05) path.

Stated bound (measured on B200): against the float64 oracle, after the full chain, z within 1e-2 of
max|z| for the 30-flow 8 kHz model and within 8e-2 for the 48-flow 22 kHz model (measured 1.5e-3 and 3.2e-2: bf16
rounding of the WaveNet activations compounds through the couplings), log-det and log_p within 2e-2 absolute; the fp32 flow variable makes forward->reverse round trips
much tighter than either pass alone is to the oracle."""
import numpy as np
import pytest
import torch

from oracle import flowavenet_oracle as O
from tests._golden import load

pytestmark = pytest.mark.gpu


def bf16_bits(t):
    return t.to(torch.bfloat16).contiguous()


@pytest.mark.parametrize("Cin,Cout,k,d,T", [(256, 256, 1, 1, 256), (256, 256, 3, 1, 300), (256, 512, 3, 3, 1000), (80, 256, 1, 1, 130),
                                            (256, 16, 1, 1, 64), (64, 256, 3, 9, 77), (16, 32, 1, 1, 5)])
def test_tcgen05_conv_vs_torch(Cin, Cout, k, d, T):
    """One implicit-GEMM launch (TMA + tcgen05.mma + TMEM epilogue) against an fp32 torch conv on the same bf16 inputs."""
    from tf_flowavenet_b200 import _lib
    B = 3
    rng = np.random.default_rng(Cin + Cout + k + d)
    x = torch.from_numpy(rng.standard_normal((B, T, Cin))).float()
    w = torch.from_numpy(rng.standard_normal((k, Cin, Cout)) / np.sqrt(k * Cin)).float()
    bias = torch.from_numpy(rng.standard_normal(Cout)).float()
    xb, wb = bf16_bits(x), bf16_bits(w)
    Cin16, Npad = (Cin + 15) // 16 * 16, (Cout + 15) // 16 * 16
    Kpad = (k * Cin16 + 63) // 64 * 64
    wp = torch.zeros(Npad, Kpad, dtype=torch.bfloat16)
    for tap in range(k):
        wp[:Cout, tap * Cin16: tap * Cin16 + Cin] = wb[tap].t()
    y = torch.empty(B, T, Cout, dtype=torch.bfloat16, device="cuda")
    xd, wd, bd = xb.cuda(), wp.cuda(), bias.cuda()
    _lib.check(_lib.lib().fwn_conv1d_bf16(_lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(y), B, T, Cin, Cout, k, d, 0, 0, None))
    torch.cuda.synchronize()
    pad = d * (k - 1) // 2
    want = torch.nn.functional.conv1d(torch.nn.functional.pad(xb.float().transpose(1, 2), (pad, pad)), wb.float().permute(2, 1, 0), bias,
                                      dilation=d).transpose(1, 2)
    err = (y.float().cpu() - want).abs().max().item()
    assert err < 2e-2 * max(1.0, want.abs().max().item()), err


@pytest.mark.parametrize("preset,B,frames", [("hparams8000", 2, 11), ("hparams", 1, 3)])
def test_mixed_full_depth_vs_oracle(preset, B, frames):
    import tf_flowavenet_b200 as P
    from tests.test_gpu_model import make_model
    ref_hp = getattr(P, preset)
    hp = O.HP(n_block=ref_hp.n_block, upsample_scales=tuple(ref_hp.upsample_scales))
    params = O.synthetic_params(hp, 77)
    x, c = O.synthetic_inputs(hp, B, frames, 78, "x")
    params = O.ddi_init(params, hp, x, c, torch.float32)
    net = make_model(hp, params, "bfloat16")
    log_p, logdet, z = net.forward(x.cuda(), c.cuda(), return_z=True)
    wlp, wld, wz = O.forward(params, hp, x, c, torch.float64)
    zerr = float((z.cpu().double() - wz).abs().max() / wz.abs().max())
    print("mixed %s: z rel-to-max err %.3e, logdet %.5f vs %.5f, log_p %.5f vs %.5f" % (preset, zerr, float(logdet), float(wld), float(log_p), float(wlp)))
    assert zerr < (1e-2 if preset == "hparams8000" else 8e-2)
    assert abs(float(logdet) - float(wld)) < 2e-2
    assert abs(float(log_p) - float(wlp)) < 2e-2
    xr = net.reverse(z, c.cuda())
    assert (xr.cpu() - x).abs().max() < 2e-2
    zin, _ = O.synthetic_inputs(hp, B, frames, 79, "z")
    got = net.reverse(zin.cuda(), c.cuda())
    want = O.reverse(params, hp, zin, c, torch.float64)
    assert (got.cpu().double() - want).abs().max() < (5e-2 if preset == "hparams8000" else 2e-1)


def test_mixed_matches_golden_small():
    hp, params, fx = load("g1_b2f2l2")
    from tests.test_gpu_model import make_model
    net = make_model(hp, params, "bfloat16")
    x, c = torch.from_numpy(fx["x"]).float().cuda(), torch.from_numpy(fx["c"]).float().cuda()
    log_p, logdet, z = net.forward(x, c, return_z=True)
    assert abs(float(log_p) - float(fx["log_p"])) < 1e-2 and abs(float(logdet) - float(fx["logdet"])) < 1e-2
    assert np.abs(z.cpu().numpy() - fx["z"]).max() < 3e-2


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_graph_replay_is_bit_identical_to_eager(dtype):
    """On a real (capturable) stream the 2nd call captures the flow chain into a CUDA graph and later calls replay it;
    every call must give bit-identical results, also after the weights change (re-prepack drops the graphs)."""
    hp, params, fx = load("g1_b2f2l2")
    from tests.test_gpu_model import make_model
    net = make_model(hp, params, dtype)
    x, c = torch.from_numpy(fx["x"]).float().cuda(), torch.from_numpy(fx["c"]).float().cuda()
    z_in = torch.from_numpy(fx["z_in"]).float().cuda()
    eager_rev = net.reverse(z_in, c)                      # legacy default stream: cannot be captured -> eager
    eager_fwd = net.forward(x, c, return_z=True)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        revs = [net.reverse(z_in, c) for _ in range(4)]   # eager, capture+launch, replay, replay
        fwds = [net.forward(x, c, return_z=True) for _ in range(4)]
    side.synchronize()
    for r in revs:
        assert torch.equal(r, eager_rev)
    for f in fwds:
        assert torch.equal(f[2], eager_fwd[2]) and float(f[0]) == float(eager_fwd[0]) and float(f[1]) == float(eager_fwd[1])
    # change one ActNorm bias: results must change accordingly (stale graphs would keep the old packed weights)
    k = "Block_0/Flow_0/ActNorm/b"
    net.load_variables({k: params[k].numpy() + 0.25})
    with torch.cuda.stream(side):
        after = [net.reverse(z_in, c) for _ in range(3)]
    side.synchronize()
    assert not torch.equal(after[0], eager_rev)
    assert torch.equal(after[0], after[1]) and torch.equal(after[1], after[2])


@pytest.mark.parametrize("B,frames", [(1, 1), (3, 1), (1, 9), (2, 33)])
def test_mixed_ragged_and_tiny_shapes(B, frames):
    """Tiles that are mostly padding (T_i < 128), batch entries that do not fill a tile, T_i not a multiple of 128:
    TMA zero fill / clipped bulk stores must reproduce the reference's per-utterance zero padding exactly."""
    import tf_flowavenet_b200 as P
    from tests.test_gpu_model import make_model
    hp = O.HP(n_block=3, n_flow=2, n_layer=2, num_mels=16, upsample_scales=(4, 2))   # hop 8 = 2^3
    params = O.synthetic_params(hp, 55)
    x, c = O.synthetic_inputs(hp, B, frames * 17, 56, "x")                            # T = 136*frames: ragged vs 128-row tiles
    net32, net16 = make_model(hp, params, "float32"), make_model(hp, params, "bfloat16")
    wlp, wld, wz = O.forward(params, hp, x, c, torch.float64)
    lp32, ld32, z32 = net32.forward(x.cuda(), c.cuda(), return_z=True)
    lp16, ld16, z16 = net16.forward(x.cuda(), c.cuda(), return_z=True)
    assert float((z32.cpu().double() - wz).abs().max() / wz.abs().max()) < 1e-4
    assert float((z16.cpu().double() - wz).abs().max() / wz.abs().max()) < 1e-2
    assert abs(float(ld32) - float(wld)) < 1e-4 * max(1.0, abs(float(wld))) and abs(float(ld16) - float(wld)) < 1e-2
    zin, _ = O.synthetic_inputs(hp, B, frames * 17, 57, "z")
    want = O.reverse(params, hp, zin, c, torch.float64)
    assert (net32.reverse(zin.cuda(), c.cuda()).cpu().double() - want).abs().max() < 1e-3
    assert (net16.reverse(zin.cuda(), c.cuda()).cpu().double() - want).abs().max() < 3e-2


@pytest.mark.parametrize("kw", [dict(n_layer=3), dict(causality=True), dict(affine=False), dict(n_layer=1), dict(n_flow=4, n_block=3)])
def test_mixed_variants_vs_oracle(kw):
    """Less-travelled graph variants on the tcgen05 path: 3 layers (skip accumulation over a middle layer, dilation 9),
    causal padding (taps at t-2d, t-d, t), additive coupling (no log-det term), a single layer, 4 flows x 3 blocks."""
    from tests.test_gpu_model import make_model
    base = dict(n_block=2, n_flow=2, n_layer=2, num_mels=8, upsample_scales=(2, 2))
    base.update(kw)
    if base["n_block"] == 3:
        base["upsample_scales"] = (4, 2)
    hp = O.HP(**base)
    params = O.synthetic_params(hp, 91)
    x, c = O.synthetic_inputs(hp, 2, 40, 92, "x")
    net = make_model(hp, params, "bfloat16")
    lp, ld, z = net.forward(x.cuda(), c.cuda(), return_z=True)
    wlp, wld, wz = O.forward(params, hp, x, c, torch.float64)
    assert float((z.cpu().double() - wz).abs().max() / wz.abs().max()) < 1e-2
    assert abs(float(ld) - float(wld)) < 1e-2 and abs(float(lp) - float(wlp)) < 1e-2
    zin, _ = O.synthetic_inputs(hp, 2, 40, 93, "z")
    want = O.reverse(params, hp, zin, c, torch.float64)
    assert (net.reverse(zin.cuda(), c.cuda()).cpu().double() - want).abs().max() < 3e-2


def test_mixed_ddi_matches_oracle():
    """ActNorm data-dependent init (train.py:221,229) through the mixed-precision pass: the statistics are taken on the fp32
    flow variable, so b / logs agree with the fp32 oracle to bf16-propagation accuracy."""
    from tests.test_gpu_model import make_model
    hp = O.HP(n_block=2, n_flow=2, n_layer=2, num_mels=8, upsample_scales=(2, 2))
    params = O.synthetic_params(hp, 94)
    x, c = O.synthetic_inputs(hp, 3, 64, 95, "x")
    net = make_model(hp, params, "bfloat16")
    lp, ld = net.initialize_actnorm(x.cuda(), c.cuda())
    want = O.ddi_init(params, hp, x, c, torch.float64)
    got = net.variables()
    for k, v in want.items():
        if "/ActNorm/" in k:
            np.testing.assert_allclose(got[k].cpu().numpy(), v.numpy(), rtol=0, atol=2e-2, err_msg=k)
    wlp, wld, _ = O.forward(want, hp, x, c, torch.float64)
    assert abs(float(ld) - float(wld)) < 2e-2 and abs(float(lp) - float(wlp)) < 2e-2
