"""Whole-pass parity of the fused C-ABI path (fwn_forward / fwn_reverse) against the oracle and against the golden
vectors produced by the reference's own Python.  Tolerances are BASELINE's: fp32 -- z and log-det 1e-4 relative,
inverted waveform 1e-3 max-abs."""
import numpy as np
import pytest
import torch

from oracle import flowavenet_oracle as O
from tests._golden import CASES, load

pytestmark = pytest.mark.gpu


def make_model(hp, params, dtype="float32"):
    import tf_flowavenet_b200 as P
    net = P.FloWaveNet(P.HParams(n_block=hp.n_block, n_flow=hp.n_flow, n_layer=hp.n_layer, num_mels=hp.num_mels, affine=hp.affine,
                                 causality=hp.causality, upsample_scales=list(hp.upsample_scales), dtype=dtype),
                       variables=P.VariableStore())
    net.load_variables({k: v.numpy() for k, v in params.items()})
    return net


def rel(a, b):
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


@pytest.mark.parametrize("case", CASES)
def test_fused_fp32_matches_golden(case):
    hp, params, fx = load(case)
    net = make_model(hp, params)
    x, c, z_in = (torch.from_numpy(fx[k]).float().cuda() for k in ("x", "c", "z_in"))
    log_p, logdet, z = net.forward(x, c, return_z=True)
    np.testing.assert_allclose(float(log_p), float(fx["log_p"]), rtol=1e-4)
    np.testing.assert_allclose(float(logdet), float(fx["logdet"]), rtol=1e-4, atol=1e-6)
    if hp.n_flow % 2 == 0:
        assert rel(z.cpu().numpy(), fx["z"]) < 1e-4
        x_rev = net.reverse(z_in, c)
        assert np.abs(x_rev.cpu().numpy() - fx["x_rev"]).max() < 1e-3
    else:  # the fused reverse refuses odd n_flow; z is defined up to the channel order -> compare as a multiset
        np.testing.assert_allclose(np.sort(z.cpu().numpy().ravel()), np.sort(fx["z"].ravel()), rtol=0, atol=1e-4)
        with pytest.raises(RuntimeError):
            net.reverse(z_in, c)
    # per-op (unfused) execution of the same graph agrees too, including odd n_flow
    lp2, ld2 = net.forward_unfused(x, c)
    np.testing.assert_allclose(float(lp2), float(fx["log_p"]), rtol=1e-4)
    np.testing.assert_allclose(float(ld2), float(fx["logdet"]), rtol=1e-4, atol=1e-6)
    xr2 = net.reverse_unfused(z_in, c)
    assert np.abs(xr2.cpu().numpy() - fx["x_rev"]).max() < 1e-3


def test_ddi_matches_golden():
    hp, params, fx = load("g1_b2f2l2")
    net = make_model(hp, params)
    x, c = torch.from_numpy(fx["x"]).float().cuda(), torch.from_numpy(fx["c"]).float().cuda()
    log_p, logdet = net.initialize_actnorm(x, c)
    np.testing.assert_allclose(float(log_p), float(fx["ddi_log_p"]), rtol=1e-4)
    np.testing.assert_allclose(float(logdet), float(fx["ddi_logdet"]), rtol=1e-4)
    v = net.variables()
    for k in fx:
        if k.startswith("ddi::"):
            np.testing.assert_allclose(v[k[5:]].cpu().numpy(), fx[k], rtol=1e-4, atol=1e-5, err_msg=k)
    lp2, ld2 = net.forward(x, c)  # second call: no re-init, same numbers
    np.testing.assert_allclose([float(lp2), float(ld2)], [float(log_p), float(logdet)], rtol=1e-6)


@pytest.mark.parametrize("preset,B,frames", [("hparams", 1, 3), ("hparams8000", 2, 11)])
def test_full_depth_fp32_vs_oracle(preset, B, frames):
    """The real 8x6x2 (and 5x6x2) models on short inputs: fp32 path vs the float64 oracle (referee)."""
    import tf_flowavenet_b200 as P
    ref_hp = getattr(P, preset)
    hp = O.HP(n_block=ref_hp.n_block, upsample_scales=tuple(ref_hp.upsample_scales))
    params = O.synthetic_params(hp, 77)
    x, c = O.synthetic_inputs(hp, B, frames, 78, "x")
    params = O.ddi_init(params, hp, x, c, torch.float32)
    net = make_model(hp, params)
    log_p, logdet, z = net.forward(x.cuda(), c.cuda(), return_z=True)
    wlp, wld, wz = O.forward(params, hp, x, c, torch.float64)
    assert rel(z.cpu().numpy(), wz.numpy()) < 1e-4
    np.testing.assert_allclose(float(logdet), float(wld), rtol=1e-4)
    np.testing.assert_allclose(float(log_p), float(wlp), rtol=1e-4)
    # round trip through the GPU path alone (size-independent property)
    xr = net.reverse(z, c.cuda())
    assert (xr.cpu() - x).abs().max() < 1e-3
    zin, _ = O.synthetic_inputs(hp, B, frames, 79, "z")
    got = net.reverse(zin.cuda(), c.cuda())
    want = O.reverse(params, hp, zin, c, torch.float64)
    assert (got.cpu().double() - want).abs().max() < 1e-3


def test_g_has_no_effect_and_is_validated():
    import tf_flowavenet_b200 as P
    hp = O.HP(n_block=2, n_flow=2, n_layer=1, num_mels=4, upsample_scales=(2, 2), gin_channels=8, n_speakers=3)
    params = O.synthetic_params(hp, 5)
    net = P.FloWaveNet(P.HParams(n_block=2, n_flow=2, n_layer=1, num_mels=4, upsample_scales=[2, 2], gin_channels=8, n_speakers=3),
                       variables=P.VariableStore())
    net.load_variables({k: v.numpy() for k, v in params.items()})
    net.load_variables({"speaker_embeddings": np.random.default_rng(0).standard_normal((3, 8))})
    x, c = O.synthetic_inputs(hp, 2, 4, 6, "x")
    with pytest.raises(ValueError, match="g is None"):  # model.py:320-321
        net.forward(x.cuda(), c.cuda())
    a = net.forward(x.cuda(), c.cuda(), g=torch.tensor([0, 2]))
    b = net.forward(x.cuda(), c.cuda(), g=torch.tensor([1, 1]))
    assert float(a[0]) == float(b[0]) and float(a[1]) == float(b[1])


def test_shape_errors():
    import tf_flowavenet_b200 as P
    hp = O.HP(n_block=2, n_flow=2, n_layer=1, num_mels=4, upsample_scales=(2, 2))
    net = make_model(hp, O.synthetic_params(hp, 1))
    x, c = O.synthetic_inputs(hp, 1, 4, 2, "x")
    with pytest.raises(ValueError):
        net.forward(x.cuda()[:, :-1], c.cuda())
    with pytest.raises(ValueError):
        net.reverse(x.cuda(), c.cuda()[:, :, :3])


def test_host_entry_points_and_chunked_synthesis():
    """fwn_reverse_host == device path; overlap-recompute chunking with receptive-field halos is exact."""
    import ctypes
    import tf_flowavenet_b200 as P
    from tf_flowavenet_b200 import _lib
    hp = O.HP(n_block=2, n_flow=2, n_layer=2, num_mels=8, upsample_scales=(2, 2))
    net = make_model(hp, O.synthetic_params(hp, 3))
    from tf_flowavenet_b200 import sharding
    halo = net.receptive_halo()
    assert halo % 4 == 0 and halo == sharding.receptive_halo(hp)
    frames = (4 * halo) // 4
    z, c = O.synthetic_inputs(hp, 1, frames, 4, "z")
    T = z.shape[1]
    full = net.reverse(z.cuda(), c.cuda())
    host = net.reverse_host(z, c)
    assert torch.equal(host, full.cpu())
    lp, ld = net.forward_host(z, c)
    a = net.forward(z.cuda(), c.cuda())
    assert abs(lp - float(a[0])) < 1e-6 and abs(ld - float(a[1])) < 1e-6
    # two chunks with halos
    mid = T // 2
    outs = []
    for lo, hi in ((0, mid), (mid, T)):
        hl, hr = (0 if lo == 0 else halo), (0 if hi == T else halo)
        ze = z[:, lo - hl:hi + hr].contiguous().cuda()
        ce = c[:, (lo - hl) // 4:(hi + hr) // 4].contiguous().cuda()
        Te = ze.shape[1]
        ws = torch.empty(_lib.lib().fwn_workspace_bytes(net._h, 1, Te), dtype=torch.uint8, device="cuda")
        xo = torch.empty(1, hi - lo, 1, device="cuda")
        _lib.check(_lib.lib().fwn_reverse_chunk(net._h, _lib.ptr(ze), _lib.ptr(ce), 1, Te, hl, hr, _lib.ptr(xo), _lib.ptr(ws), ws.numel(), None))
        outs.append(xo)
    got = torch.cat(outs, 1)
    assert (got - full).abs().max() < 1e-5
    # the Python wrapper of the same entry point, as used by FloWaveNet.reverse_sharded
    again = net.reverse_chunk(z[:, :mid + halo].contiguous().cuda(), c[:, :(mid + halo) // 4].contiguous().cuda(), 0, halo)
    assert torch.equal(again, outs[0])


def test_synthesis_caller_npy_to_wav(tmp_path):
    """synthesize.py:23-49 counterpart: .npy mels in, wav files out; the wav equals the device path's waveform (16-bit PCM)."""
    import types
    import tf_flowavenet_b200 as P
    from tf_flowavenet_b200 import synthesize as S
    hp = O.HP(n_block=2, n_flow=2, n_layer=2, num_mels=8, upsample_scales=(2, 2))
    params = O.synthetic_params(hp, 9)
    hparams = P.HParams(n_block=2, n_flow=2, n_layer=2, num_mels=8, upsample_scales=[2, 2], sample_rate=8000, temp=0.7)
    mels, out = tmp_path / "mels", tmp_path / "out"
    mels.mkdir()
    rng = np.random.default_rng(1)
    for i, frames in enumerate((16, 23)):  # ragged lengths
        np.save(mels / ("utt%d.npy" % i), rng.uniform(0, 1, (frames, 8)).astype(np.float32))
    np.savez(tmp_path / "w.npz", **{k: v.numpy() for k, v in params.items()})
    names = S.synthesize(types.SimpleNamespace(weights=str(tmp_path / "w.npz"), mels_dir=str(mels), output_dir=str(out), seed=5), hparams)
    assert names == ["utt0.npy", "utt1.npy"]
    # reproduce utterance 0 through the device API with the same z stream
    model = S.get_model(hparams, {k: v.numpy() for k, v in params.items()})
    r2 = np.random.default_rng(5)
    mel0 = np.load(mels / "utt0.npy")
    z0 = (r2.standard_normal((1, 16 * 4, 1)) * 0.7).astype(np.float32)
    want = model.reverse(torch.from_numpy(z0).cuda(), torch.from_numpy(mel0[None]).cuda()).cpu().numpy().reshape(-1)
    got, sr = S.read_wav(str(out / "utt0.wav"))
    assert sr == 8000 and got.shape == want.shape
    np.testing.assert_allclose(got, np.clip(want, -1, 1), atol=1.0 / 32767 + 1e-6)
    got1, _ = S.read_wav(str(out / "utt1.wav"))
    assert got1.shape == (23 * 4,)
    # and against the oracle on the same z (fp32 tolerance of BASELINE: 1e-3 max-abs)
    ref = O.reverse(params, hp, torch.from_numpy(z0), torch.from_numpy(mel0[None]), torch.float64).numpy().reshape(-1)
    assert np.abs(want - ref).max() < 1e-3
    # --saved_dir, the reference's own flag (synthesize.py:28-34): a tf.train.Saver checkpoint directory read without TensorFlow --
    # variables under vocoder/FloWaveNet/ next to Adam slots and global_step -- gives the same wavs as the .npz
    from tf_flowavenet_b200 import checkpoint as C
    ckpt = {"vocoder/FloWaveNet/" + k: v.numpy() for k, v in params.items()}
    ckpt.update({"vocoder/FloWaveNet/" + k + "/Adam": np.zeros_like(v.numpy()) for k, v in params.items()})
    ckpt["global_step"] = np.array(500000, dtype=np.int64)
    C.write_checkpoint(str(tmp_path / "logs" / "model.ckpt-500000"), ckpt)
    out2 = tmp_path / "out2"
    S.synthesize(types.SimpleNamespace(saved_dir=str(tmp_path / "logs"), weights=None, mels_dir=str(mels), output_dir=str(out2), seed=5), hparams)
    for n in ("utt0.wav", "utt1.wav"):
        assert open(out / n, "rb").read() == open(out2 / n, "rb").read()
