"""The training-step oracle itself (CPU): autograd gradients against central finite differences of the pinned forward graph,
the learning-rate schedule (train.py:15-20), clip_by_global_norm (train.py:27-32), the tower average (utils.py:34-60), Adam."""
import numpy as np
import torch

from oracle import flowavenet_oracle as O
from oracle import flowavenet_train_oracle as TO


def tiny():
    hp = O.HP(n_block=2, n_flow=2, n_layer=2, num_mels=4, upsample_scales=(2, 2))
    p = O.synthetic_params(hp, seed=3, dtype=torch.float64)
    x, c = O.synthetic_inputs(hp, 2, 4, 11, "forward")
    p = O.ddi_init(p, hp, x, c, torch.float64)
    return hp, p, x, c


def test_autograd_matches_finite_differences():
    hp, p, x, c = tiny()
    loss, log_p, logdet, grads = TO.loss_and_grads(p, hp, x, c)
    assert abs(loss + log_p + logdet) < 1e-12
    rng = np.random.default_rng(0)

    def f(pp):
        lp, ld, _ = O.forward(pp, hp, x, c, torch.float64)
        return float(-(lp + ld))

    probes = ["Block_0/Flow_0/ActNorm/logs", "Block_1/Flow_1/ActNorm/b", "conv2d_transpose/kernel", "conv2d_transpose_1/wn/g",
              "Block_0/Flow_1/AffineCoupling/WaveNet/Conv_front/conv1d/kernel",
              "Block_1/Flow_0/AffineCoupling/WaveNet/ResBlock_0_0/Conv_gate/conv1d/wn/g",
              "Block_1/Flow_0/AffineCoupling/WaveNet/ResBlock_0_1/conv1d_1/kernel",
              "Block_0/Flow_0/AffineCoupling/WaveNet/ResBlock_0_0/conv1d_2/bias",
              "Block_1/Flow_1/AffineCoupling/WaveNet/ZeroConv1d/scale", "Block_0/Flow_0/AffineCoupling/WaveNet/ZeroConv1d/conv1d/kernel"]
    for name in probes:
        flat = p[name].reshape(-1)
        for idx in rng.choice(flat.numel(), size=min(3, flat.numel()), replace=False):
            h = 1e-6
            pp = {k: v.clone() for k, v in p.items()}
            pp[name].reshape(-1)[idx] += h
            fp = f(pp)
            pp[name].reshape(-1)[idx] -= 2 * h
            fm = f(pp)
            fd = (fp - fm) / (2 * h)
            ag = float(grads[name].reshape(-1)[idx])
            assert abs(fd - ag) <= 1e-6 + 1e-5 * abs(ag), (name, idx, fd, ag)
    # dead variables: the last layer's residual conv (modules.py:170-176) gets no gradient
    assert float(grads["Block_0/Flow_0/AffineCoupling/WaveNet/ResBlock_0_1/conv1d_2/kernel"].abs().max()) == 0.0


def test_learning_rate_schedule():
    assert TO.learning_rate(0) == 0.001 and TO.learning_rate(199999) == 0.001
    assert TO.learning_rate(200000) == 0.0005 and TO.learning_rate(399999) == 0.0005
    assert TO.learning_rate(400000) == 0.00025
    assert abs(TO.learning_rate(600000) - 0.001 / 6) < 1e-15


def test_clip_average_adam():
    g1 = {"a": torch.tensor([3.0, 0.0]), "b": torch.tensor([[4.0]])}
    clipped, n = TO.clip_by_global_norm(g1, 1.0)
    assert abs(n - 5.0) < 1e-12 and abs(TO.global_norm(clipped) - 1.0) < 1e-6
    small, n2 = TO.clip_by_global_norm({"a": torch.tensor([0.3])}, 1.0)
    assert abs(float(small["a"]) - 0.3) < 1e-7  # below the threshold: unchanged
    avg = TO.average_gradients([g1, {"a": torch.tensor([1.0, 2.0]), "b": torch.tensor([[0.0]])}])
    assert torch.allclose(avg["a"], torch.tensor([2.0, 1.0])) and torch.allclose(avg["b"], torch.tensor([[2.0]]))
    # first Adam step moves every coordinate by lr * sign(g) (bias-corrected)
    p = {"a": torch.zeros(2, dtype=torch.float64)}
    m = {"a": torch.zeros(2, dtype=torch.float64)}
    v = {"a": torch.zeros(2, dtype=torch.float64)}
    new = TO.adam_step(p, {"a": torch.tensor([0.5, -2.0], dtype=torch.float64)}, m, v, 1e-3, 1)
    assert torch.allclose(new["a"], torch.tensor([-1e-3, 1e-3], dtype=torch.float64), rtol=1e-6)


import pytest  # noqa: E402

from tests._golden import load  # noqa: E402

GRAD_CASES = ["g1_b2f2l2", "g2_b3f2l1", "g4_causal", "g5_additive", "g6_l3"]


def load_grads(case):
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", case + "_grads.npz"))


@pytest.mark.parametrize("case", GRAD_CASES)
def test_oracle_gradients_match_reference_fixture(case):
    """tests/golden/*_grads.npz hold tf.gradients(loss, trainable_variables) of the reference's own model.py (train.py:59-63) run on
    the TF shim (make_golden_grads.py): this pins the training oracle the same way the forward fixtures pin the forward oracle."""
    hp, params, fx = load(case)
    gx = load_grads(case)
    loss, log_p, logdet, grads = TO.loss_and_grads(params, hp, torch.from_numpy(fx["x"]), torch.from_numpy(fx["c"]))
    assert abs(loss - float(gx["loss"])) < 1e-6 * max(1.0, abs(loss))  # the reference returns float32 scalars (model.py:345-347)
    # the reference keeps ActNorm variables and the returned scalars in float32 (model.py:24-27,345-347): agreement to ~1e-7
    assert abs(TO.global_norm(grads) - float(gx["global_norm"])) < 1e-6 * float(gx["global_norm"])
    for k, g in grads.items():
        flat = g.reshape(-1).numpy()
        n = float(gx["norm::" + k])
        assert abs(float(np.sqrt((flat * flat).sum())) - n) <= 1e-6 * max(n, 1e-12), k
        np.testing.assert_allclose(flat[gx["idx::" + k]], gx["val::" + k], rtol=1e-5, atol=1e-6 * max(n, 1e-12), err_msg=k)
