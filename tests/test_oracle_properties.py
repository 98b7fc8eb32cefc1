"""Known-answer properties that hold for ANY correct implementation of the path."""
import math

import numpy as np
import torch

from oracle import flowavenet_oracle as O


def _small(n_block=2, n_flow=2, n_layer=2, mels=4, scales=(2, 2), **kw):
    return O.HP(n_block=n_block, n_flow=n_flow, n_layer=n_layer, num_mels=mels, upsample_scales=scales, **kw)


def test_squeeze_index_law_and_inverse():
    x = torch.arange(2 * 8 * 3, dtype=torch.float64).reshape(2, 8, 3)
    s = O.squeeze(x)
    for b in range(2):
        for t in range(4):
            for c in range(3):
                for k in range(2):
                    assert s[b, t, 2 * c + k] == x[b, 2 * t + k, c]  # SURVEY F8 / model.py:226-228
    assert torch.equal(O.unsqueeze(s), x)


def test_n_fold_squeeze_is_bit_reversal():
    n, T = 3, 32
    x = torch.arange(T, dtype=torch.float64).reshape(1, T, 1)
    s = x
    for _ in range(n):
        s = O.squeeze(s)
    for ch in range(2 ** n):
        off = int(format(ch, "0%db" % n)[::-1], 2)
        assert s[0, 1, ch] == 2 ** n + off


def test_reverse_inverts_forward():
    hp = _small()
    p = O.synthetic_params(hp, 3, torch.float64)
    x, c = O.synthetic_inputs(hp, 2, 8, 5, "x")
    _, _, z = O.forward(p, hp, x, c)
    xr = O.reverse(p, hp, z, c)
    np.testing.assert_allclose(xr.numpy(), x.double().numpy(), atol=1e-12)


def test_logdet_is_slogdet_over_T():
    hp = _small(n_block=2, n_flow=2, n_layer=1, scales=(2, 2))
    p = O.synthetic_params(hp, 4, torch.float64)
    x, c = O.synthetic_inputs(hp, 1, 4, 6, "x")  # T = 16
    T = x.shape[1]

    def f(v):
        return O.forward(p, hp, v.reshape(1, T, 1), c)[2].reshape(T)

    J = torch.autograd.functional.jacobian(f, x.double().reshape(T))
    _, logabs = torch.linalg.slogdet(J)
    _, logdet, _ = O.forward(p, hp, x, c)
    np.testing.assert_allclose(float(logdet), float(logabs) / T, rtol=1e-10)


def test_zero_init_coupling_is_identity():
    hp = _small()
    p = O.synthetic_params(hp, 5, torch.float64)
    for k in p:
        if "ZeroConv1d" in k:
            p[k] = torch.zeros_like(p[k])  # the reference's real init (modules.py:46-49)
    x, c = O.synthetic_inputs(hp, 1, 8, 7, "x")
    _, logdet, _ = O.forward(p, hp, x, c)
    want = sum(float((3.0 * v).mean()) for k, v in p.items() if k.endswith("/ActNorm/logs"))
    np.testing.assert_allclose(float(logdet), want, rtol=1e-12)


def test_ddi_gives_zero_mean_unit_rms():
    x = torch.randn(3, 50, 4, dtype=torch.float64) * 2.5 + 1.0
    b, logs = O.actnorm_ddi(x)
    y, _ = O.actnorm_forward({"a/b": b, "a/logs": logs}, "a", x)
    np.testing.assert_allclose(y.mean(dim=(0, 1)).numpy(), 0.0, atol=1e-12)
    np.testing.assert_allclose((y ** 2).mean(dim=(0, 1)).numpy(), 1.0, rtol=1e-6)


def test_conv_matches_explicit_tap_sum():
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.standard_normal((2, 11, 3)))
    p = {"c/conv1d/kernel": torch.from_numpy(rng.standard_normal((3, 3, 5))), "c/conv1d/wn/g": torch.from_numpy(rng.uniform(0.5, 1.5, 5)),
         "c/conv1d/bias": torch.from_numpy(rng.standard_normal(5))}
    for d in (1, 3):
        y = O.conv(p, "c", x, 3, d, False)
        v = p["c/conv1d/kernel"]
        w = v / torch.sqrt((v * v).sum(dim=(0, 1), keepdim=True)) * p["c/conv1d/wn/g"]
        want = torch.zeros(2, 11, 5, dtype=torch.float64)
        for t in range(11):
            for k in range(3):
                tt = t + (k - 1) * d
                if 0 <= tt < 11:
                    want[:, t] += x[:, tt] @ w[k]
        want += p["c/conv1d/bias"]
        np.testing.assert_allclose(y.numpy(), want.numpy(), atol=1e-12)


def test_upsample_is_gradient_of_same_conv():
    """[TF] conv2d_transpose(SAME) := d/d(input) of the SAME forward conv; check the oracle's closed form."""
    rng = np.random.default_rng(1)
    for s in (2, 8, 12, 16):
        x = torch.from_numpy(rng.standard_normal((1, 5, 7)))
        w = torch.from_numpy(rng.standard_normal((2 * s, 3)))
        got = O.conv2d_transpose_same(x, w, s)
        img = torch.zeros(1, 1, 5 * s, 7, dtype=torch.float64, requires_grad=True)
        ph = s  # (H-1)*s + 2s - s*H
        y = torch.nn.functional.conv2d(torch.nn.functional.pad(img, (1, 1, ph // 2, ph - ph // 2)), w[None, None], stride=(s, 1))
        (g,) = torch.autograd.grad(y, img, grad_outputs=x[:, None])
        np.testing.assert_allclose(got.numpy(), g[:, 0].numpy(), atol=1e-12)


def test_flop_count_matches_survey():
    """8 655 870 MAC/sample for hparams.py, 7 724 296 for hparams8000.py."""
    def macs(n_block, scales):
        tot = 0.0
        for i in range(n_block):
            per_step = 3 * 2 ** i * 256 + 2 * (2 * 3 * 256 * 256 + 2 * 256 * 256) + 256 * 256 + 256 * 2 ** (i + 1) + 2 * 2 * 80 * 2 ** i * 256
            tot += 6 * per_step / 2 ** (i + 1)
        return tot + sum(2 * s * 3 / 1 for s in scales[-1:]) * 0  # upsampler counted separately
    assert round(macs(8, (16, 16))) == 6689280 + 1966080
    assert round(macs(5, (8, 12))) == 6494976 + 1228800
