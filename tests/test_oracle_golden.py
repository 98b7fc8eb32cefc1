"""Oracle vs golden vectors produced by the reference's own Python (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import flowavenet_oracle as O
from tests._golden import CASES, load


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_run(case):
    hp, params, fx = load(case)
    x, c, z_in = (torch.from_numpy(fx[k]) for k in ("x", "c", "z_in"))
    log_p, logdet, z = O.forward(params, hp, x, c, torch.float64)
    # 3e-8: the reference multiplies by a float32 sqrt(0.5) constant (modules.py:128); the oracle uses a double
    np.testing.assert_allclose(float(log_p), float(fx["log_p"]), rtol=2e-7)
    np.testing.assert_allclose(float(logdet), float(fx["logdet"]), rtol=2e-6, atol=1e-9)
    np.testing.assert_allclose(z.numpy(), fx["z"], rtol=0, atol=2e-6)
    x_rev = O.reverse(params, hp, z_in, c, torch.float64)
    np.testing.assert_allclose(x_rev.numpy(), fx["x_rev"], rtol=0, atol=2e-6)
    c_up = O.upsample(params, c.double(), hp.upsample_scales)
    np.testing.assert_allclose(c_up.numpy(), fx["c_up"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("case,count", [("g1_b2f2l2", None), ("g7_hparams8000", 47 * 30 + 6), ("g8_hparams", 47 * 48 + 6)])
def test_variable_names_match_reference_scoping(case, count):
    """Names created by the reference's own scoping code; the full-depth cases pin the variable count of the shipped
    configurations (hparams.py: 47 per flow x 48 flows + 6 upsampler = 2 262)."""
    hp, params, fx = load(case)
    names = str(fx["names"]).split("\n")
    assert sorted(O.param_shapes(hp)) == names
    if count is not None:
        assert len(names) == count


def test_ddi_matches_reference_run():
    hp, params, fx = load("g1_b2f2l2")
    x, c = torch.from_numpy(fx["x"]), torch.from_numpy(fx["c"])
    pp = O.ddi_init(params, hp, x, c, torch.float64)
    for k in fx:
        if k.startswith("ddi::"):
            # the reference stores ActNorm variables in float32 (model.py:33 dtype default)
            np.testing.assert_allclose(pp[k[5:]].numpy(), fx[k], rtol=1e-6, atol=2e-7, err_msg=k)
    log_p, logdet, _ = O.forward(pp, hp, x, c, torch.float64)
    np.testing.assert_allclose(float(log_p), float(fx["ddi_log_p"]), rtol=1e-6)
    np.testing.assert_allclose(float(logdet), float(fx["ddi_logdet"]), rtol=1e-6)


def test_fp32_oracle_close_to_fp64():
    hp, params, fx = load("g1_b2f2l2")
    x, c = torch.from_numpy(fx["x"]), torch.from_numpy(fx["c"])
    lp, ld, z = O.forward(params, hp, x, c, torch.float32)
    np.testing.assert_allclose(z.double().numpy(), fx["z"], rtol=0, atol=5e-5)
    np.testing.assert_allclose(float(ld), float(fx["logdet"]), rtol=1e-4)
