"""Per-op parity: each C-ABI op (through the reference-named Python classes) vs the CPU oracle, fp32.
Tolerances are stated per test; integer/index ops (squeeze, change_order) are bit-exact."""
import numpy as np
import pytest
import torch

from oracle import flowavenet_oracle as O

pytestmark = pytest.mark.gpu


def dev(t):
    return t.float().cuda().contiguous()


def rnd(*shape, seed=0, scale=1.0):
    return torch.from_numpy(np.random.default_rng(seed).standard_normal(shape) * scale).float()


@pytest.mark.parametrize("B,T,C", [(1, 2, 1), (2, 64, 1), (3, 34, 2), (3, 34, 4), (5, 6, 3), (1, 4096, 80), (2, 10, 160), (1, 8, 6)])
def test_squeeze_unsqueeze_bit_exact(B, T, C):
    from tf_flowavenet_b200.model import _squeeze
    x = rnd(B, T, C, seed=1)
    y = _squeeze(dev(x))
    assert torch.equal(y.cpu(), O.squeeze(x))
    assert torch.equal(_squeeze(y, True).cpu(), x)


@pytest.mark.parametrize("rows,C", [(1, 2), (77, 6), (1000, 256), (5, 20480)])
def test_change_order_bit_exact(rows, C):
    import tf_flowavenet_b200 as P
    x, c = rnd(1, rows, C, seed=2), rnd(1, rows, 2 * C, seed=3)
    xs, cs, g = P.change_order(dev(x), dev(c))
    wx, wc = O.change_order(x, c)
    assert torch.equal(xs.cpu(), wx) and torch.equal(cs.cpu(), wc) and g is None


@pytest.mark.parametrize("B,T,C", [(2, 50, 2), (1, 333, 4), (2, 64, 256), (1, 7, 6)])
def test_actnorm_fwd_rev_ddi(B, T, C):
    import tf_flowavenet_b200 as P
    store = P.VariableStore()
    x = rnd(B, T, C, seed=4, scale=2.0) + 0.7
    p = {"ActNorm/b": rnd(1, 1, C, seed=5, scale=0.2), "ActNorm/logs": rnd(1, 1, C, seed=6, scale=0.1)}
    store.update({k: dev(v) for k, v in p.items()})
    an = P.ActNorm(C, variables=store)
    y, obj = an(dev(x))
    wy, wobj = O.actnorm_forward(p, "ActNorm", x)
    np.testing.assert_allclose(y.cpu().numpy(), wy.numpy(), rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(float(obj), float(wobj), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(an.reverse(y).cpu().numpy(), x.numpy(), rtol=1e-5, atol=1e-5)
    # data-dependent init: zero mean / unit RMS output, values equal to the oracle's
    an2 = P.ActNorm(C, init=True, variables=P.VariableStore(), scope="A2")
    y2, _ = an2(dev(x))
    wb, wl = O.actnorm_ddi(x.double())
    np.testing.assert_allclose(an2._store["A2/b"].cpu().numpy(), wb.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(an2._store["A2/logs"].cpu().numpy(), wl.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(y2.cpu().double().mean(dim=(0, 1)).numpy(), 0, atol=1e-5)
    np.testing.assert_allclose((y2.cpu().double() ** 2).mean(dim=(0, 1)).numpy(), 1, rtol=1e-4)


@pytest.mark.parametrize("Cin,Cout,k,d,causal", [(1, 256, 3, 1, False), (256, 256, 3, 3, False), (256, 256, 3, 9, False),
                                                 (80, 256, 1, 1, False), (7, 12, 3, 1, True), (256, 256, 3, 3, True), (128, 256, 3, 1, False)])
def test_conv_matches_oracle(Cin, Cout, k, d, causal):
    import tf_flowavenet_b200 as P
    B, T = 2, 150
    rng = np.random.default_rng(7)
    p = {"C/conv1d/kernel": torch.from_numpy(rng.standard_normal((k, Cin, Cout))).float(),
         "C/conv1d/wn/g": torch.from_numpy(rng.uniform(0.5, 1.5, Cout)).float(), "C/conv1d/bias": torch.from_numpy(rng.standard_normal(Cout)).float()}
    store = P.VariableStore({n: dev(v) for n, v in p.items()})
    x = rnd(B, T, Cin, seed=8)
    y = P.Conv(Cin, Cout, k, d, causal, scope="C", variables=store)(dev(x))
    want = O.conv({n: v.double() for n, v in p.items()}, "C", x.double(), k, d, causal)
    np.testing.assert_allclose(y.cpu().numpy(), want.numpy(), rtol=1e-4, atol=2e-5)


def test_conv_ragged_and_tiny_shapes():
    """T not a multiple of the 128-row tile, T smaller than the dilation, single row."""
    import tf_flowavenet_b200 as P
    rng = np.random.default_rng(9)
    for T in (1, 2, 5, 127, 129, 257):
        p = {"C/conv1d/kernel": torch.from_numpy(rng.standard_normal((3, 5, 9))).float(), "C/conv1d/wn/g": torch.ones(9),
             "C/conv1d/bias": torch.zeros(9)}
        x = rnd(3, T, 5, seed=T)
        y = P.Conv(5, 9, 3, 3, False, scope="C", variables=P.VariableStore({n: dev(v) for n, v in p.items()}))(dev(x))
        want = O.conv({n: v.double() for n, v in p.items()}, "C", x.double(), 3, 3, False)
        np.testing.assert_allclose(y.cpu().numpy(), want.numpy(), rtol=1e-4, atol=2e-5)


def _wavenet_params(hp, seed):
    allp = O.synthetic_params(hp, seed)
    pre = "Block_0/Flow_0/AffineCoupling/"
    return {k[len(pre):]: v for k, v in allp.items() if k.startswith(pre + "WaveNet")}, allp


@pytest.mark.parametrize("n_layer,causal", [(2, False), (3, False), (2, True)])
def test_wavenet_and_resblock(n_layer, causal):
    import tf_flowavenet_b200 as P
    hp = O.HP(n_block=1, n_flow=1, n_layer=n_layer, num_mels=8, upsample_scales=(2,), causality=causal)
    p, _ = _wavenet_params(hp, 21)
    store = P.VariableStore({k: dev(v) for k, v in p.items()})
    x, c = rnd(2, 70, 1, seed=22), rnd(2, 70, 8, seed=23)
    net = P.WaveNet(1, 2, 1, n_layer, 256, 256, 256, 3, 8, causal, variables=store)
    y = net(dev(x), dev(c), None)
    want = O.wavenet({k: v.double() for k, v in p.items()}, "WaveNet", x.double(), c.double(), n_layer, causal)
    np.testing.assert_allclose(y.cpu().numpy(), want.numpy(), rtol=2e-4, atol=2e-5)
    h = rnd(2, 70, 256, seed=24)
    out, skip = net._res_blocks[1](dev(h), dev(c))
    wo, ws = O.resblock({k: v.double() for k, v in p.items()}, "WaveNet/ResBlock_0_1", h.double(), c.double(), 3, causal)
    np.testing.assert_allclose(out.cpu().numpy(), wo.numpy(), rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(skip.cpu().numpy(), ws.numpy(), rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("affine", [True, False])
def test_coupling_flow_block(affine):
    import tf_flowavenet_b200 as P
    hp = O.HP(n_block=1, n_flow=2, n_layer=2, num_mels=4, upsample_scales=(2,), affine=affine)
    p = O.synthetic_params(hp, 31)
    store = P.VariableStore({k: dev(v) for k, v in p.items()})
    x, c = rnd(2, 40, 1, seed=32), rnd(2, 40, 4, seed=33)
    blk = P.Block(1, 4, 2, 2, init=False, affine=affine, scope="Block_0", variables=store)
    out, cc, g, logdet = blk(dev(x), dev(c))
    wo, wc, wl = O.block_forward({k: v.double() for k, v in p.items()}, "Block_0", x.double(), c.double(), 2, 2, affine)
    np.testing.assert_allclose(out.cpu().numpy(), wo.numpy(), rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(cc.cpu().numpy(), wc.numpy(), rtol=0, atol=0)
    np.testing.assert_allclose(float(logdet), float(wl), rtol=1e-4, atol=1e-6)
    xr, cr, _ = blk.reverse(out, cc)
    np.testing.assert_allclose(xr.cpu().numpy(), x.numpy(), rtol=0, atol=1e-4)
    assert torch.equal(cr.cpu(), c)


@pytest.mark.parametrize("s", [2, 8, 12, 16])
def test_upsample_stage(s):
    import tf_flowavenet_b200 as P
    hp = O.HP(n_block=1, n_flow=2, n_layer=1, num_mels=80, upsample_scales=(s,))
    p = O.synthetic_params(hp, 41)
    net = P.FloWaveNet(P.HParams(n_block=1, n_flow=2, n_layer=1, num_mels=80, upsample_scales=[s]))
    net.load_variables({k: v.numpy() for k, v in p.items()})
    c = torch.from_numpy(np.random.default_rng(42).uniform(0, 1, (2, 9, 80))).float()
    y = net.upsample(dev(c))
    want = O.upsample({k: v.double() for k, v in p.items()}, c.double(), (s,))
    np.testing.assert_allclose(y.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-6)


def test_log_p_and_empty_inputs():
    from tf_flowavenet_b200 import _lib
    z = rnd(3, 1000, 1, seed=51)
    out = torch.empty((), device="cuda")
    zc = dev(z)
    _lib.check(_lib.lib().fwn_log_p(_lib.ptr(zc), _lib.ptr(out), zc.numel(), None))
    np.testing.assert_allclose(float(out), float((0.5 * (-O.LOG_2PI - z.double() ** 2)).mean()), rtol=1e-6)
    # zero-length elementwise calls are no-ops, not errors
    e = torch.empty(0, device="cuda")
    _lib.check(_lib.lib().fwn_add(_lib.ptr(e), _lib.ptr(e), _lib.ptr(e), 0, 0, None))
    with pytest.raises(RuntimeError):
        _lib.check(_lib.lib().fwn_squeeze(_lib.ptr(zc), _lib.ptr(zc), 1, 3, 1, None))  # odd T
