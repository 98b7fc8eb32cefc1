"""Parity of the CUDA path AT THE SHAPES BASELINE.json BENCHMARKS, through the C ABI, every test printing the error it
measured:

  C2      hparams.py (8x6x2, 80 mels) forward log-likelihood, 8 x 16 128 samples, fp32      vs the float64 oracle (z, log-det, log_p <= 1e-4)
  C3      hparams8000.py inverse synthesis, a 2 x 834-frame slice of the 32 x 834 batch, mixed  vs the float64 oracle (stated bound)
  deep    hparams.py with 1 024 frames (T = 262 144): block 7 has M = 1 024 rows against K_c = 10 240 (eight row tiles of the
          largest contraction of the model), fp32 (1e-4 / 1e-3) and mixed (stated bound), both directions
  C4      the 60 s utterance (T = 1 323 008) cut into 8 time chunks with receptive-field halos through fwn_reverse_chunk, each chunk
          bit-equal to the unsharded pass (model.py:350-396 semantics across chunk borders: modules.py:27 zero padding only at
          the true utterance edges)

Reference graph: model.py:317-347 (forward), 350-396 (reverse).  Tolerances are BASELINE.json's for fp32 (z / log-det 1e-4 relative,
waveform 1e-3 max-abs); the mixed-precision bounds are the measured ones plus margin and are written next to each assert.
"""
import math

import numpy as np
import pytest
import torch

from oracle import flowavenet_oracle as O

pytestmark = pytest.mark.gpu


def _model(preset, dtype):
    import tf_flowavenet_b200 as P
    ref_hp = getattr(P, preset)
    hp = O.HP(n_block=ref_hp.n_block, upsample_scales=tuple(ref_hp.upsample_scales))
    net = P.FloWaveNet(P.HParams(**{**ref_hp.values(), "dtype": dtype}), variables=P.VariableStore())
    return hp, net


def _load(net, params):
    net.load_variables({k: v.numpy() for k, v in params.items()})


def snr_db(got, want):
    err = float(((got - want) ** 2).sum())
    return 10.0 * math.log10(float((want ** 2).sum()) / max(err, 1e-300))


def relmax(got, want):
    return float((got - want).abs().max() / want.abs().max())


_CACHE = {}


def _case(preset, B, frames, seed):
    """Seeded weights (ActNorm by DDI on the case's own batch, train.py:221) + inputs + the float64 oracle's outputs, computed once."""
    key = (preset, B, frames, seed)
    if key not in _CACHE:
        import tf_flowavenet_b200 as P
        ref_hp = getattr(P, preset)
        hp = O.HP(n_block=ref_hp.n_block, upsample_scales=tuple(ref_hp.upsample_scales))
        params = O.synthetic_params(hp, seed)
        x, c = O.synthetic_inputs(hp, B, frames, seed + 1, "x")
        zin, _ = O.synthetic_inputs(hp, B, frames, seed + 2, "z")
        torch.set_num_threads(max(torch.get_num_threads(), 8))
        with torch.no_grad():
            params = O.ddi_init(params, hp, x, c, torch.float32)
            wlp, wld, wz = O.forward(params, hp, x, c, torch.float64)
            wx = O.reverse(params, hp, zin, c, torch.float64)
        _CACHE.clear()   # one case resident at a time (the 8-block parameter set is 0.7 GB)
        _CACHE[key] = (hp, params, x, c, zin, float(wlp), float(wld), wz, wx)
    return _CACHE[key]


# ---------------------------------------------------------------------------------------------------------------- C2
def test_c2_forward_fp32_full_shape():
    """BASELINE config 2 at its full shape: 8 x 16 128 samples (63 frames), fp32 parity mode."""
    hp, params, x, c, zin, wlp, wld, wz, wx = _case("hparams", 8, 63, 1234 + 2)
    _, net = _model("hparams", "float32")
    _load(net, params)
    lp, ld, z = net.forward(x.cuda(), c.cuda(), return_z=True)
    ez = relmax(z.cpu().double(), wz)
    print("C2 fp32 forward 8x16128: z rel-to-max %.3e, logdet %.7f vs %.7f (rel %.2e), log_p %.7f vs %.7f (rel %.2e)" %
          (ez, float(ld), wld, abs(float(ld) - wld) / abs(wld), float(lp), wlp, abs(float(lp) - wlp) / abs(wlp)))
    assert ez < 1e-4
    np.testing.assert_allclose(float(ld), wld, rtol=1e-4)
    np.testing.assert_allclose(float(lp), wlp, rtol=1e-4)
    got = net.reverse(zin.cuda(), c.cuda())
    ex = float((got.cpu().double() - wx).abs().max())
    print("C2 shape fp32 inverse: waveform max-abs err %.3e, SNR %.1f dB" % (ex, snr_db(got.cpu().double(), wx)))
    assert ex < 1e-3


# ---------------------------------------------------------------------------------------------------------------- C3
# Measured on B200 (profiles/r2_parity_shapes.md) -> stated bound (about 3x margin):
#   30-flow 8 kHz model : bf16 z 3.5e-3, waveform 2.1e-3 / 59 dB;  fp16 z 3.6e-4, waveform 2.5e-4 / 77 dB
#   48-flow 22 kHz model: bf16 z 4.2e-3, waveform 2.8e-3 / 57 dB;  fp16 z 4.9e-4, waveform 3.3e-4 / 75 dB
MIXED_BOUNDS = {
    # (preset, dtype): (z rel-to-max, waveform max-abs, waveform SNR dB)
    ("hparams8000", "bfloat16"): (1e-2, 1e-2, 50.0),
    ("hparams8000", "float16"): (1.5e-3, 1e-3, 68.0),
    ("hparams", "bfloat16"): (1.5e-2, 1e-2, 48.0),
    ("hparams", "float16"): (2e-3, 1.5e-3, 66.0),
}


def _mixed_dtypes():
    from tf_flowavenet_b200 import _lib
    return ["bfloat16", "float16"] if hasattr(_lib, "FWN_MIXED_FP16") else ["bfloat16"]


@pytest.mark.parametrize("dtype", ["bfloat16", "float16"])
def test_c3_slice_mixed_reverse(dtype):
    """BASELINE config 3: two of the 32 utterances at the full 834 frames (T = 80 064), mixed precision, inverse synthesis."""
    if dtype not in _mixed_dtypes():
        pytest.skip("fp16 operand mode not built")
    hp, params, x, c, zin, wlp, wld, wz, wx = _case("hparams8000", 2, 834, 1234 + 3)
    _, net = _model("hparams8000", dtype)
    _load(net, params)
    zb, xb, sb = MIXED_BOUNDS[("hparams8000", dtype)]
    got = net.reverse(zin.cuda(), c.cuda()).cpu().double()
    ex, snr = float((got - wx).abs().max()), snr_db(got, wx)
    lp, ld, z = net.forward(x.cuda(), c.cuda(), return_z=True)
    ez = relmax(z.cpu().double(), wz)
    print("C3 slice %s 2x80064: waveform max-abs err %.3e (|x|max %.2f), SNR %.1f dB; forward z rel-to-max %.3e, logdet err %.2e, log_p err %.2e" %
          (dtype, ex, float(wx.abs().max()), snr, ez, abs(float(ld) - wld), abs(float(lp) - wlp)))
    assert ex < xb and snr > sb and ez < zb
    assert abs(float(ld) - wld) < 2e-2 and abs(float(lp) - wlp) < 2e-2


# ---------------------------------------------------------------------------------------------------------------- deep
@pytest.mark.parametrize("dtype", ["float32", "bfloat16", "float16"])
def test_full_depth_1024_frames(dtype):
    """hparams.py at 1 024 frames: every block's GEMMs run over several row tiles, block 7 with M = 1 024 x K_c = 10 240."""
    if dtype != "float32" and dtype not in _mixed_dtypes():
        pytest.skip("fp16 operand mode not built")
    hp, params, x, c, zin, wlp, wld, wz, wx = _case("hparams", 1, 1024, 4321)
    _, net = _model("hparams", dtype)
    _load(net, params)
    lp, ld, z = net.forward(x.cuda(), c.cuda(), return_z=True)
    ez = relmax(z.cpu().double(), wz)
    got = net.reverse(zin.cuda(), c.cuda()).cpu().double()
    ex, snr = float((got - wx).abs().max()), snr_db(got, wx)
    rt = float((net.reverse(z, c.cuda()).cpu() - x).abs().max())
    print("deep %s 1x262144: z rel-to-max %.3e, logdet err %.2e, log_p err %.2e; waveform max-abs err %.3e (|x|max %.2f), SNR %.1f dB; "
          "round trip %.2e" % (dtype, ez, abs(float(ld) - wld), abs(float(lp) - wlp), ex, float(wx.abs().max()), snr, rt))
    if dtype == "float32":
        assert ez < 1e-4 and ex < 1e-3 and rt < 1e-3
        np.testing.assert_allclose(float(ld), wld, rtol=1e-4)
        np.testing.assert_allclose(float(lp), wlp, rtol=1e-4)
    else:
        zb, xb, sb = MIXED_BOUNDS[("hparams", dtype)]
        assert ez < zb and ex < xb and snr > sb and rt < 2e-2
        assert abs(float(ld) - wld) < 2e-2 and abs(float(lp) - wlp) < 2e-2


# ---------------------------------------------------------------------------------------------------------------- C4
@pytest.mark.parametrize("dtype", ["bfloat16"])
def test_c4_eight_chunks_equal_unsharded(dtype):
    """BASELINE config 4 at full size: the 60 s utterance in 8 time chunks with receptive-field halos == the unsharded pass, bit for bit
    (every output row of the implicit GEMMs is computed from the same operands in the same order wherever its tile starts)."""
    import tf_flowavenet_b200 as P
    from tf_flowavenet_b200 import sharding
    hp, net = _model("hparams", dtype)
    _load(net, O.synthetic_params(hp, 1234))
    xi, ci = O.synthetic_inputs(hp, 2, 64, 99, "x")
    net.initialize_actnorm(xi.cuda(), ci.cuda())
    frames, world = 5168, 8
    z, c = O.synthetic_inputs(hp, 1, frames, 1234 + 4, "z")
    T, hop = z.shape[1], hp.hop
    halo = net.receptive_halo()
    assert halo == sharding.receptive_halo(P.hparams) == 15616
    full = net.reverse(z.cuda(), c.cuda())
    assert torch.isfinite(full).all()
    worst = 0.0
    for (lo, hi) in sharding.chunk_bounds(T, world, hop):
        hl, hr = (0 if lo == 0 else halo), (0 if hi == T else halo)
        ze = z[:, lo - hl:hi + hr].contiguous().cuda()
        ce = c[:, (lo - hl) // hop:(hi + hr) // hop].contiguous().cuda()
        got = net.reverse_chunk(ze, ce, hl, hr)
        worst = max(worst, float((got - full[:, lo:hi]).abs().max()))
        assert torch.equal(got, full[:, lo:hi]), "chunk [%d,%d) differs from the unsharded pass by %.3e" % (lo, hi, worst)
    # a halo 1 024 samples short of the receptive field (15 300 + one hop) is NOT exact: the bound is tight, not just sufficient
    lo, hi = sharding.chunk_bounds(T, world, hop)[3]
    short = halo - 1024
    ze = z[:, lo - short:hi + short].contiguous().cuda()
    ce = c[:, (lo - short) // hop:(hi + short) // hop].contiguous().cuda()
    assert not torch.equal(net.reverse_chunk(ze, ce, short, short), full[:, lo:hi])
    # size-independent property at the full size: forward(reverse(z)) == z
    _, _, zz = net.forward(full, c.cuda(), return_z=True)
    rt = float((zz.cpu() - z).abs().max())
    print("C4 %s 1x%d: 8 chunks vs unsharded max-abs diff %.1e (bit-equal); forward(reverse(z)) - z max-abs %.2e" % (dtype, T, worst, rt))
    assert rt < 2e-2


# ---------------------------------------------------------------------------------------------------------------- C3, full size
def test_c3_full_batch_properties():
    """BASELINE config 3 at the full size bench.py times (32 x 80 064 samples, bf16 operands; multi-wave launches of the fused layer /
    tail kernels, 68 row-tile pairs per CTA pair): size-independent properties, since the oracle cannot run this shape in seconds.
      * forward(reverse(z)) == z (model.py:317-396): the fp32 flow variable makes the round trip much tighter than either pass is
        to the oracle; bound 2e-2 max-abs on N(0, 0.7) inputs (measured, printed)
      * batch independence: utterances 0 and 31 of the batch == the same utterances run alone, bit for bit (every output row of the
        implicit GEMMs is computed from the same operands in the same order wherever its tile sits in the batch)
      * the fused kernels == one launch per GEMM at this size, bit for bit."""
    hp, net = _model("hparams8000", "bfloat16")
    _load(net, O.synthetic_params(hp, 1234))
    xi, ci = O.synthetic_inputs(hp, 2, 64, 99, "x")
    net.initialize_actnorm(xi.cuda(), ci.cuda())
    B, frames = 32, 834
    z, c = O.synthetic_inputs(hp, B, frames, 1234 + 3, "z")
    zd, cd = z.cuda(), c.cuda()
    x = net.reverse(zd, cd)
    assert torch.isfinite(x).all()
    _, _, zz = net.forward(x, cd, return_z=True)
    rt = float((zz - zd).abs().max())
    for b in (0, B - 1):
        alone = net.reverse(zd[b:b + 1].contiguous(), cd[b:b + 1].contiguous())
        assert torch.equal(alone[0], x[b]), "utterance %d depends on its batch" % b
    net.set_layer_fusion(0)
    x2 = net.reverse(zd, cd)
    net.set_layer_fusion(-1)
    assert torch.equal(x, x2), "fused kernels differ from the one-launch-per-GEMM path by %.3e" % float((x - x2).abs().max())
    print("C3 full batch %dx%d bf16: forward(reverse(z)) - z max-abs %.2e; batch-independent and fused == unfused (bit-equal)" % (B, z.shape[1], rt))
    assert rt < 2e-2
