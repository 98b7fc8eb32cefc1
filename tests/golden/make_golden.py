"""Generate tests/golden/*.npz by running the REFERENCE'S OWN Python on the eager TF shim.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

What runs: /root/reference/model.py, modules.py, convolutional.py -- imported unmodified -- with
``oracle/tf_shim`` standing in for ``tensorflow`` (TF 1.12 is not installable here).  Weights are the
seeded synthetic set from ``oracle.flowavenet_oracle.synthetic_params`` injected BY NAME, so the
run also proves that the oracle's variable-name schema is the one the reference's scoping creates.
Fixtures hold inputs + outputs + a weight checksum (weights themselves are re-derived from the seed).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import tensorflow as tf  # noqa: E402  (the shim)
from oracle import flowavenet_oracle as O  # noqa: E402

CASES = {
    # name: (hp kwargs, B, n_frames, seed)
    "g1_b2f2l2": (dict(n_block=2, n_flow=2, n_layer=2, num_mels=8, upsample_scales=(2, 2)), 2, 8, 11),
    "g2_b3f2l1": (dict(n_block=3, n_flow=2, n_layer=1, num_mels=4, upsample_scales=(4, 2)), 1, 5, 12),
    "g3_oddflow": (dict(n_block=2, n_flow=3, n_layer=2, num_mels=4, upsample_scales=(2, 2)), 2, 6, 13),
    "g4_causal": (dict(n_block=2, n_flow=2, n_layer=2, num_mels=4, upsample_scales=(2, 2), causality=True), 1, 10, 14),
    "g5_additive": (dict(n_block=2, n_flow=2, n_layer=2, num_mels=4, upsample_scales=(2, 2), affine=False), 2, 6, 15),
    "g6_l3": (dict(n_block=1, n_flow=2, n_layer=3, num_mels=6, upsample_scales=(2,)), 1, 24, 16),
    # full depth, 80 mels: the real K_c ladder (80 .. 10240) and variable count of hparams8000.py (5x6x2, 1 416 names) and
    # hparams.py (8x6x2, 2 262 names, model.py:293-299) on a few frames
    "g7_hparams8000": (dict(n_block=5, upsample_scales=(8, 12)), 1, 3, 17),
    "g8_hparams": (dict(n_block=8, upsample_scales=(16, 16)), 1, 2, 18),
}


def run_reference(hp, params, x, c, z_in, dtype, ddi=False):
    tf.reset_default_graph()
    for m in ("model", "modules", "convolutional", "utils"):
        sys.modules.pop(m, None)
    import model as ref_model  # /root/reference/model.py

    tf.set_presets({"FloWaveNet/" + k: v.numpy() for k, v in params.items()})
    hparams = tf.contrib.training.HParams(
        n_block=hp.n_block, n_flow=hp.n_flow, n_layer=hp.n_layer, num_mels=hp.num_mels, affine=hp.affine,
        causality=hp.causality, upsample_scales=list(hp.upsample_scales), gin_channels=-1, n_speakers=7, dtype=dtype)
    init = tf.convert_to_tensor(True) if ddi else False  # a tensor, not a Python bool (model.py:34-39)
    net = ref_model.FloWaveNet(hparams, init=init, scope="FloWaveNet")
    xt, ct = tf.convert_to_tensor(x.to(dtype)), tf.convert_to_tensor(c.to(dtype))
    log_p, logdet = net.forward(xt, ct)
    out = {"log_p": float(log_p), "logdet": float(logdet)}
    if ddi:
        # freeze: later calls must not re-init
        out["vars"] = {k[len("FloWaveNet/"):]: v.detach().clone() for k, v in tf.global_variables_dict().items()}
        rep = tf.shim_report()
        assert not rep["created_without_preset"], rep
        return out
    # z: replay FloWaveNet.forward's own loop (model.py:326-340) with the reference's blocks
    with tf.variable_scope(net._vs, auxiliary_name_scope=False):
        o, cc, g = xt, net.upsample(ct), None
        for blk in net._blocks:
            o, cc, g, _ = blk(o, cc, g)
    out["z_sq"] = o.detach().as_subclass(torch.Tensor).clone()
    out["c_up"] = net.upsample(ct).detach().as_subclass(torch.Tensor).clone()
    out["x_rev"] = net.reverse(tf.convert_to_tensor(z_in.to(dtype)), ct).detach().as_subclass(torch.Tensor).clone()
    rep = tf.shim_report()
    assert not rep["created_without_preset"], rep
    assert not rep["unused_presets"], rep
    out["names"] = sorted(k[len("FloWaveNet/"):] for k in tf.global_variables_dict())
    return out


def main():
    only = set(sys.argv[1:])   # optional: regenerate just the named cases
    for name, (kw, B, nf, seed) in CASES.items():
        if only and name not in only:
            continue
        hp = O.HP(**kw)
        params = O.synthetic_params(hp, seed=seed, dtype=torch.float64)
        x, c = O.synthetic_inputs(hp, B, nf, seed + 100, "x")
        z_in, _ = O.synthetic_inputs(hp, B, nf, seed + 200, "z")
        ref = run_reference(hp, params, x, c, z_in, tf.float64)
        z_flat = ref["z_sq"]
        for _ in range(hp.n_block):
            z_flat = O.unsqueeze(z_flat)
        chk = np.array([sum(float(v.sum()) for v in params.values()), sum(float((v * v).sum()) for v in params.values())])
        fx = dict(hp=np.array(repr(kw)), B=B, n_frames=nf, seed=seed, x=x.numpy(), c=c.numpy(), z_in=z_in.numpy(),
                  log_p=ref["log_p"], logdet=ref["logdet"], z=z_flat.numpy(), z_sq=ref["z_sq"].numpy(),
                  c_up=ref["c_up"].numpy(), x_rev=ref["x_rev"].numpy(), weight_checksum=chk,
                  names=np.array("\n".join(ref["names"])))
        if name == "g1_b2f2l2":  # data-dependent init pass (train.py:221: feed init=True)
            d = run_reference(hp, params, x, c, z_in, tf.float64, ddi=True)
            fx["ddi_log_p"], fx["ddi_logdet"] = d["log_p"], d["logdet"]
            for k, v in d["vars"].items():
                if "/ActNorm/" in k:
                    fx["ddi::" + k] = v.as_subclass(torch.Tensor).numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **fx)
        print(name, "log_p=%.9f logdet=%.9f |x_rev|max=%.4f" % (ref["log_p"], ref["logdet"], float(ref["x_rev"].abs().max())))


if __name__ == "__main__":
    main()
