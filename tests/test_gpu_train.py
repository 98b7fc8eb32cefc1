"""Training step through the C ABI (fwn_train_enable / fwn_loss_and_grads / fwn_apply_gradients) against the
training oracle: gradients of -(log_p + logdet) w.r.t. EVERY variable, the device-side re-pack, clip + Adam trajectories.
Tolerance (stated): every variable's gradient within 5e-4 of its own max-abs (fp32 accumulation, float64 referee)."""
import numpy as np
import pytest
import torch

from oracle import flowavenet_oracle as O
from oracle import flowavenet_train_oracle as TO
from tests._golden import load
from tests.test_gpu_model import make_model

pytestmark = pytest.mark.gpu

GRAD_CASES = ["g1_b2f2l2", "g2_b3f2l1", "g4_causal", "g5_additive", "g6_l3"]


def check_grads(got, ref, tol=5e-4):
    gmax = max(float(r.abs().max()) for r in ref.values())
    worst = ("", 0.0)
    for k, r in ref.items():
        g = got[k].double().cpu()
        r = r.double()
        scale = max(float(r.abs().max()), 1e-6 * gmax)
        err = float((g - r).abs().max()) / scale
        if err > worst[1]:
            worst = (k, err)
    assert worst[1] < tol, worst


@pytest.mark.parametrize("terms", [3, 6])
@pytest.mark.parametrize("case", GRAD_CASES)
def test_gradients_match_oracle(case, terms):
    """terms = bf16 products per fp32 product on the tensor cores: 3 (training default, ~2^-16 per product) and 6 (fp32 accuracy)."""
    import tf_flowavenet_b200.train as T
    hp, params, fx = load(case)
    net = make_model(hp, params)
    tr = T.Trainer(net, split_terms=terms)
    x, c = torch.from_numpy(fx["x"]).float().cuda(), torch.from_numpy(fx["c"]).float().cuda()
    log_p, logdet = tr.loss_and_grads(x, c)
    np.testing.assert_allclose(float(log_p), float(fx["log_p"]), rtol=1e-4)
    np.testing.assert_allclose(float(logdet), float(fx["logdet"]), rtol=1e-4, atol=1e-6)
    _, _, _, ref = TO.loss_and_grads(params, hp, torch.from_numpy(fx["x"]), torch.from_numpy(fx["c"]))
    check_grads(tr.gradients(), ref, tol=5e-4 if terms == 3 else 2e-4)
    # deterministic up to atomics: a second call agrees to fp32 rounding
    g1 = tr.grads.clone()
    tr.loss_and_grads(x, c)
    assert float((tr.grads - g1).abs().max()) <= 1e-4 * float(g1.abs().max())


@pytest.mark.parametrize("case", GRAD_CASES)
def test_gradients_match_reference_fixture(case):
    """Against tf.gradients of the reference's own model.py run on the TF shim (tests/golden/make_golden_grads.py, train.py:59-63):
    per-variable gradient norms and probed entries."""
    import os
    import tf_flowavenet_b200.train as T
    hp, params, fx = load(case)
    gx = np.load(os.path.join(os.path.dirname(__file__), "golden", case + "_grads.npz"))
    tr = T.Trainer(make_model(hp, params), split_terms=6)
    x, c = torch.from_numpy(fx["x"]).float().cuda(), torch.from_numpy(fx["c"]).float().cuda()
    log_p, logdet = tr.loss_and_grads(x, c)
    assert abs(float(-(log_p + logdet)) - float(gx["loss"])) < 1e-4 * max(1.0, abs(float(gx["loss"])))
    norm = torch.zeros(1, device="cuda")
    got = tr.gradients()
    gmax = max(float(gx["norm::" + k]) for k in got)
    sq = 0.0
    for k, g in got.items():
        flat = g.reshape(-1).double().cpu().numpy()
        n = float(gx["norm::" + k])
        sq += float((flat * flat).sum())
        assert abs(float(np.sqrt((flat * flat).sum())) - n) <= 2e-4 * max(n, 1e-6 * gmax), k
        np.testing.assert_allclose(flat[gx["idx::" + k]], gx["val::" + k], rtol=0, atol=2e-4 * max(n, 1e-6 * gmax), err_msg=k)
    assert abs(np.sqrt(sq) - float(gx["global_norm"])) < 1e-4 * float(gx["global_norm"])


def test_device_repack_equals_host_prepack():
    """fwn_train_enable re-derives every operand on the device; the forward result must stay on the golden values, and a
    forward in inference mode after it (same handle) as well."""
    import tf_flowavenet_b200.train as T
    hp, params, fx = load("g1_b2f2l2")
    net = make_model(hp, params)
    x, c = torch.from_numpy(fx["x"]).float().cuda(), torch.from_numpy(fx["c"]).float().cuda()
    T.Trainer(net)
    log_p, logdet, z = net.forward(x, c, return_z=True)
    np.testing.assert_allclose(float(log_p), float(fx["log_p"]), rtol=1e-4)
    np.testing.assert_allclose(float(logdet), float(fx["logdet"]), rtol=1e-4, atol=1e-6)
    assert float(np.abs(z.cpu().numpy() - fx["z"]).max() / np.abs(fx["z"]).max()) < 1e-4


def test_training_trajectory_matches_oracle():
    """Three optimizer steps (clip_by_global_norm 1 + Adam 1e-3, train.py:15-32,76-81) against the float64 oracle."""
    import tf_flowavenet_b200.train as T
    hp, params, fx = load("g1_b2f2l2")
    net = make_model(hp, params)
    tr = T.Trainer(net)
    x, c = torch.from_numpy(fx["x"]).float().cuda(), torch.from_numpy(fx["c"]).float().cuda()
    p = {k: v.double() for k, v in params.items()}
    m = {k: torch.zeros_like(v) for k, v in p.items()}
    v = {k: torch.zeros_like(v) for k, v in p.items()}
    p0 = {k: t.clone() for k, t in p.items()}
    for step in range(3):
        info = tr.train_step(x, c)
        p, ref = TO.train_step(p, hp, [(torch.from_numpy(fx["x"]), torch.from_numpy(fx["c"]))], m, v, step)
        assert abs(float(info["loss"]) - ref["loss"]) < 1e-4 * max(1.0, abs(ref["loss"])), (step, float(info["loss"]), ref["loss"])
        np.testing.assert_allclose(float(info["grad_global_norm"]), ref["grad_global_norm"], rtol=1e-3)
        assert info["learning_rate"] == ref["lr"]
    got = tr.variables()
    lr = 1e-3
    dev, tot = 0.0, 0
    for k in p:
        d_ref = (p[k] - p0[k])
        d_got = got[k].double().cpu() - p0[k]
        # Adam normalises each coordinate: coordinates with |g| ~ eps are ill-conditioned, so bound the mean deviation
        dev += float((d_got - d_ref).abs().sum())
        tot += d_ref.numel()
        assert float((d_got - d_ref).abs().max()) <= 3 * lr * 3 + 1e-7, k
    assert dev / tot < 0.02 * lr, dev / tot
    # the loss went down
    lp, ld = net.forward(x, c)
    assert float(-(lp + ld)) < float(-(fx["log_p"] + fx["logdet"]))


def test_init_step_runs_ddi_then_trains():
    """train.py:221,229: the first sess.run feeds init=True -- ActNorm statistics from the batch, then a normal update."""
    import tf_flowavenet_b200.train as T
    hp, params, fx = load("g1_b2f2l2")
    net = make_model(hp, params)
    tr = T.Trainer(net)
    x, c = torch.from_numpy(fx["x"]).float().cuda(), torch.from_numpy(fx["c"]).float().cuda()
    info = tr.train_step(x, c, init=True)
    np.testing.assert_allclose(float(info["log_p"]), float(fx["ddi_log_p"]), rtol=1e-4)
    np.testing.assert_allclose(float(info["logdet"]), float(fx["ddi_logdet"]), rtol=1e-4)
    pd = O.ddi_init({k: v.double() for k, v in params.items()}, hp, torch.from_numpy(fx["x"]), torch.from_numpy(fx["c"]), torch.float64)
    _, _, _, ref = TO.loss_and_grads(pd, hp, torch.from_numpy(fx["x"]), torch.from_numpy(fx["c"]))
    # gradient of the init step is taken at the DDI values
    tr2 = T.Trainer(make_model(hp, {k: v.float() for k, v in pd.items()}))
    tr2.loss_and_grads(x, c)
    check_grads(tr2.gradients(), ref)


def test_training_needs_fp32_master_variables():
    import tf_flowavenet_b200 as P
    import tf_flowavenet_b200.train as T
    hp, params, fx = load("g1_b2f2l2")
    with pytest.raises(ValueError):
        T.Trainer(make_model(hp, params, dtype="bfloat16"))   # compute dtype is a Trainer argument, the model keeps fp32 variables
    with pytest.raises(ValueError):
        T.Trainer(make_model(hp, params), compute_dtype="int8")


def test_two_stream_backward_equals_single_stream_at_full_depth():
    """The backward pass runs weight gradients and the conditioning gradient on a side stream (event-ordered, double-buffered
    scratch).  Same gradients as the single-stream order on the full hparams.py model at the C5 shape, up to fp32 atomics order."""
    import os
    import tf_flowavenet_b200 as P
    import tf_flowavenet_b200.train as T
    from tf_flowavenet_b200.synthetic import synthetic_inputs, synthetic_params
    net = P.FloWaveNet(P.HParams(**{**P.hparams.values(), "dtype": "float32"}), variables=P.VariableStore())
    net.load_variables(synthetic_params(net.variable_shapes(), seed=5))
    x, c = synthetic_inputs(256, 80, 8, 25, 6, "x")
    x, c = torch.from_numpy(x).cuda(), torch.from_numpy(c).cuda()
    net.initialize_actnorm(x, c)
    tr = T.Trainer(net)
    res = {}
    try:
        for mode in ("1", "0", "0"):
            os.environ["FWN_TRAIN_STREAMS"] = mode
            tr.loss_and_grads(x, c)
            torch.cuda.synchronize()
            res.setdefault(mode, []).append(tr.grads[:tr._np].clone())
    finally:
        os.environ.pop("FWN_TRAIN_STREAMS", None)
    single, dual_a, dual_b = res["1"][0], res["0"][0], res["0"][1]
    scale = float(single.abs().max())
    assert float((dual_a - single).abs().max()) < 1e-4 * scale
    assert float((dual_b - dual_a).abs().max()) < 1e-4 * scale
    assert abs(float(dual_a.double().norm()) - float(single.double().norm())) < 1e-5 * float(single.double().norm())


def check_grads_global(got, ref, tol_max, tol_norm):
    """max |error| over ALL variables relative to the largest gradient entry of the model, and the relative L2 error of the whole vector."""
    gmax = max(float(r.abs().max()) for r in ref.values())
    worst = max((float((got[k].double().cpu() - r.double()).abs().max()) / gmax, k) for k, r in ref.items())
    assert worst[0] < tol_max, worst
    num = sum(float(((got[k].double().cpu() - r.double()) ** 2).sum()) for k, r in ref.items()) ** 0.5
    den = sum(float((r.double() ** 2).sum()) for r in ref.values()) ** 0.5
    assert num / den < tol_norm, num / den


@pytest.mark.parametrize("B,n_frames", [(1, 160), (3, 100)])
def test_gradients_match_oracle_multi_tile(B, n_frames):
    """Rows spanning several (partial) 128-row tiles and 64-step wgrad chunks, several utterances, and both conditioning paths:
    T_i = 320/200 in block 0 (projection inside the gate GEMM) and 160/100 in block 1 (projection computed ahead on the side stream).

    Tolerance.  On these longer sequences some gradients (upsampler g, front-conv kernels) are small differences of large sums, and
    the backward quantities are ill-conditioned in the forward rounding: swapping only the FORWARD gate/res-skip GEMMs between
    the split engine and the CUDA-core engine (outputs agree to 5e-6 relative -- the tensor cores' accumulator truncates where
    FFMA rounds) moves d log_s by 1e-3 (tools/debug_tape.py); a float32 run of the oracle itself stays within 3e-5 of float64
    here, so the gap is the tensor cores' truncating accumulator, not fp32 conditioning.  Until the engine promotes
    its partial sums (two-level accumulation) these cases are bounded against the model's gradient scale:
    6 terms -- every entry within 1e-3 of the largest gradient entry (measured 2.3e-4), whole vector within 1e-3 relative L2;
    3 terms (training default of bench.py) -- 5e-3 / 5e-3 (measured 1.3e-3).  The well-conditioned small cases above are held
    to per-variable bounds."""
    import tf_flowavenet_b200.train as T
    hp = O.HP(n_block=2, n_flow=2, n_layer=2, num_mels=8, upsample_scales=(2, 2))
    params = O.synthetic_params(hp, seed=21, dtype=torch.float64)
    x, c = O.synthetic_inputs(hp, B, n_frames, 22, "x")
    params = O.ddi_init(params, hp, x, c, torch.float64)
    loss, _, _, ref = TO.loss_and_grads(params, hp, x, c)
    for terms, exact in ((6, True), (6, False), (3, False)):
        tr = T.Trainer(make_model(hp, params), split_terms=terms, exact_forward=exact)
        log_p, logdet = tr.loss_and_grads(x.float().cuda(), c.float().cuda())
        assert abs(float(-(log_p + logdet)) - loss) < 1e-4 * max(1.0, abs(loss))
        if exact:
            # the parity setting (Trainer default): forward GEMMs round to nearest (CUDA-core FFMA), backward GEMMs on the split engine
            # -> the PER-VARIABLE bound of the small cases holds on long sequences too
            check_grads(tr.gradients(), ref, tol=2e-4)
        elif terms == 6:
            check_grads_global(tr.gradients(), ref, 1e-3, 1e-3)
        else:
            check_grads_global(tr.gradients(), ref, 5e-3, 5e-3)


def test_gradient_buckets_tile_the_flat_vector():
    """fwn_grad_bucket_range: production order (last block first), contiguous ranges tiling [0, param_floats); equal to the Python
    twin used by the gloo test."""
    import tf_flowavenet_b200 as P
    import tf_flowavenet_b200.train as T
    for gin in (-1, 4):
        hp = P.HParams(n_block=3, n_flow=2, n_layer=1, num_mels=4, upsample_scales=[4, 2], gin_channels=gin, n_speakers=3, dtype="float32")
        net = P.FloWaveNet(hp, variables=P.VariableStore())
        net.init_variables(seed=1, zero_init_coupling=False)
        tr = T.Trainer(net)
        assert tr.buckets == T.bucket_ranges(net.variable_shapes(), 3)
        cover = sorted(tr.buckets)
        assert cover[0][0] == 0 and cover[-1][0] + cover[-1][1] == tr.param_floats()
        assert all(a[0] + a[1] == b[0] for a, b in zip(cover, cover[1:]))
        assert tr.buckets[0][0] > tr.buckets[1][0] > tr.buckets[2][0] > 0 and tr.buckets[3][0] == 0


def test_variables_after_training_and_partial_reload():
    """after optimizer steps the HANDLE owns the live variables.  FloWaveNet.variables() must return the trained values (what
    a checkpoint for synthesize.py would hold), and load_variables() of ONE variable must neither resurrect the initial values of
    the others nor leave the dgrad operands stale."""
    import tf_flowavenet_b200.train as T
    hp, params, fx = load("g1_b2f2l2")
    net = make_model(hp, params)
    tr = T.Trainer(net)
    x, c = torch.from_numpy(fx["x"]).float().cuda(), torch.from_numpy(fx["c"]).float().cuda()
    for _ in range(2):
        tr.train_step(x, c)
    live, via_model = tr.variables(), net.variables()
    moved = 0
    for k, v in live.items():
        assert torch.equal(via_model[k], v), k
        moved += int(not torch.equal(v.cpu().double(), params[k].double().reshape(v.shape)))
    assert moved > len(live) // 2
    # overwrite one variable; everything else keeps its trained value, gradients match the oracle at the combined point
    k = "Block_1/Flow_0/AffineCoupling/WaveNet/Conv_final/conv1d/bias"
    new = live[k] + 0.05
    net.load_variables({k: new.cpu().numpy()})
    tr.loss_and_grads(x, c)
    point = {n: v.double().cpu() for n, v in live.items()}
    point[k] = new.double().cpu()
    _, _, _, ref = TO.loss_and_grads(point, hp, torch.from_numpy(fx["x"]), torch.from_numpy(fx["c"]))
    check_grads(tr.gradients(), ref, tol=2e-4)
    after = tr.variables()
    assert torch.equal(after[k], new) and all(torch.equal(after[n], live[n]) for n in live if n != k)


def test_dataset_feeds_trainer(tmp_path):
    """dataset.py -> train.py wiring: TFRecord clips -> hop-aligned crops -> pinned per-tower batches -> train_step (with speaker ids)."""
    import tf_flowavenet_b200 as P
    import tf_flowavenet_b200.train as T
    from tf_flowavenet_b200 import dataset as D
    hp = P.HParams(n_block=2, n_flow=2, n_layer=2, num_mels=8, upsample_scales=[2, 2], hop_size=4, max_time_steps=64, batch_size=2,
                   gin_channels=4, n_speakers=3, dtype="float32")
    path = str(tmp_path / "train.tfrecord")
    rng = np.random.default_rng(0)
    with D.TFRecordWriter(path) as w:
        for i in range(5):
            frames = int(rng.integers(10, 40))
            audio = (0.5 * np.sin(np.arange(frames * 4) * 0.1 * (i + 1)) + 0.05 * rng.standard_normal(frames * 4)).astype(np.float32)
            audio, mel = D.adjust_time_resolution(audio, rng.random((frames, 8)).astype(np.float32), hp)
            w.write(D.make_example(audio, mel, i % 3))
    net = P.FloWaveNet(hp, variables=P.VariableStore())
    net.init_variables(seed=1, zero_init_coupling=False)
    tr = T.Trainer(net)
    it = iter(D.Dataset(path, hp, num_towers=1, seed=2))
    losses = []
    for step in range(4):
        mel, audio, spk = next(it)[0]
        assert tuple(audio.shape) == (2, 64, 1) and tuple(mel.shape) == (2, 16, 8)
        info = tr.train_step(audio.cuda(non_blocking=True), mel.cuda(non_blocking=True), spk.cuda(non_blocking=True), init=(step == 0))
        losses.append(float(info["loss"]))
    assert all(np.isfinite(losses)) and tr.global_step == 4


def test_tf_checkpoint_save_restore(tmp_path):
    """Trainer.save_checkpoint writes tf.train.Saver's TensorBundle format (train.py:190,252); restoring it into a fresh trainer
    continues the run, and the model variables in it are what synthesize --saved_dir loads."""
    import tf_flowavenet_b200.train as T
    from tf_flowavenet_b200 import checkpoint as C
    hp, params, fx = load("g1_b2f2l2")
    x, c = torch.from_numpy(fx["x"]).float().cuda(), torch.from_numpy(fx["c"]).float().cuda()
    tr = T.Trainer(make_model(hp, params))
    for _ in range(2):
        tr.train_step(x, c)
    prefix = tr.save_checkpoint(str(tmp_path / "logs" / "model.ckpt-2"))
    ck = C.load_checkpoint(prefix)
    assert int(ck["global_step"]) == 2 and "vocoder/FloWaveNet/Block_0/Flow_0/ActNorm/logs/Adam_1" in ck
    live = tr.variables()
    for k, v in C.flowavenet_variables(C.latest_checkpoint(str(tmp_path / "logs"))).items():
        np.testing.assert_array_equal(v, live[k].cpu().numpy())
    a = tr.train_step(x, c)
    tr2 = T.Trainer(make_model(hp, params))
    tr2.restore_checkpoint(prefix)
    assert tr2.global_step == 2
    b = tr2.train_step(x, c)
    assert abs(float(a["loss"]) - float(b["loss"])) < 1e-5 * max(1.0, abs(float(a["loss"])))


def test_checkpoint_resume_continues_the_same_trajectory():
    """train.py:190,199-210: saving variables + Adam slots + step and restoring them into a fresh trainer continues bit-compatibly
    (up to the order of fp32 atomics in the weight gradients)."""
    import tf_flowavenet_b200.train as T
    hp, params, fx = load("g1_b2f2l2")
    x, c = torch.from_numpy(fx["x"]).float().cuda(), torch.from_numpy(fx["c"]).float().cuda()
    tr = T.Trainer(make_model(hp, params))
    for _ in range(2):
        tr.train_step(x, c)
    state = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in tr.state_dict().items()}
    for _ in range(2):
        a = tr.train_step(x, c)
    tr2 = T.Trainer(make_model(hp, params))       # fresh handle, initial variables, zero moments
    tr2.load_state_dict(state)
    assert tr2.global_step == 2
    for _ in range(2):
        b = tr2.train_step(x, c)
    assert abs(float(a["loss"]) - float(b["loss"])) < 1e-5 * max(1.0, abs(float(a["loss"])))
    assert b["learning_rate"] == a["learning_rate"] and b["global_step"] == a["global_step"] == 4
    va, vb = tr.state_dict(), tr2.state_dict()
    for k in ("variables", "adam_m", "adam_v"):
        assert float((va[k] - vb[k]).abs().max()) <= 1e-5 * max(float(va[k].abs().max()), 1e-12), k
