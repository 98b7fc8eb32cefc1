"""CPU oracle for the FloWaveNet flow pass -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``tf_flowavenet_b200``) never imports it and has no CPU fallback.

What it is: a restatement, in PyTorch-CPU (float64 or float32), of the reference's
TensorFlow-1.12 graph for the hot path:

  * model.py        (ActNorm 7-105, AffineCoupling 108-164, change_order 166-174,
                     Flow 176-205, Block 207-280, FloWaveNet 282-404)
  * modules.py      (Conv 6-36, ZeroConv1d 39-59, ResBlock 62-131, WaveNet 134-189)
  * convolutional.py (weight-normed Conv1D.build 53-109, Conv2DTranspose.build 155-201)

The arithmetic of the reference lives in a third-party dependency that is absent from
/root/reference: ``tensorflow-gpu==1.12`` (requirements.txt:3).  TF 1.12 cannot be installed
here (no cp312 wheel, no network), so the TF primitives are restated from their published
semantics (marked [TF] below).

Pinning status: the reference ships no tests and no golden vectors, so parity against a real
TF-1.12 run is UNPINNED.  What pins this oracle instead:
  1. tests/golden/*.npz -- produced by running the reference's OWN UNMODIFIED Python
     (model.py / modules.py / convolutional.py imported from /root/reference) on top of
     ``oracle/tf_shim`` (an eager stand-in for the handful of TF ops it calls).  That pins
     all composition logic (squeeze order, change_order, flow order, log-det accounting,
     variable naming); only the [TF] primitives are shared restatements.
  2. Known-answer properties that hold for any correct implementation
     (tests/test_oracle_properties.py): reverse(forward) == id, logdet == slogdet(J)/T,
     zero-init coupling == identity, DDI statistics, squeeze index law, g has no effect.
  3. An independent einsum restatement of the dilated conv and a gradient-of-SAME-conv
     definition of the transposed conv (tests/test_oracle_properties.py).

Layout: channels-last [B, T, C] everywhere, exactly like the reference.
Parameters: a dict name -> torch tensor using the reference's variable names relative to the
model scope, e.g. ``Block_0/Flow_3/AffineCoupling/WaveNet/ResBlock_0_1/Conv_gate/conv1d/kernel``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

LOG_2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------------------
# [TF] primitives
# --------------------------------------------------------------------------------------
def l2_normalize(v: torch.Tensor, axes: Sequence[int]) -> torch.Tensor:
    """[TF] nn_impl.l2_normalize: x * rsqrt(max(sum(x^2, axes), 1e-12)) (convolutional.py:80,186)."""
    ss = (v * v).sum(dim=tuple(axes), keepdim=True)
    return v * torch.rsqrt(torch.clamp(ss, min=1e-12))


def weight_normed_kernel(p: Params, prefix: str, axes=(0, 1)) -> torch.Tensor:
    """convolutional.py:73-83 -- kernel = l2_normalize(v, [0,1]) * g (per output channel)."""
    v = p[prefix + "/kernel"]
    gname = prefix + "/wn/g"
    if gname in p:
        return l2_normalize(v, axes) * p[gname]
    return v


def conv1d_valid(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], dilation: int) -> torch.Tensor:
    """[TF] nn_ops.Convolution VALID, channels-last, cross-correlation (convolutional.py:102-108).

    x [B,T,Cin], w [k,Cin,Cout] -> [B, T - d(k-1), Cout];  y[t,o] = sum_k sum_i x[t + k d, i] w[k,i,o] + b[o]
    """
    y = F.conv1d(x.transpose(1, 2), w.permute(2, 1, 0), b, dilation=dilation)
    return y.transpose(1, 2)


def conv(p: Params, prefix: str, x: torch.Tensor, kernel_size: int, dilation: int, causal: bool) -> torch.Tensor:
    """modules.py Conv.forward 24-33: zero pad both sides then VALID dilated conv."""
    pad = dilation * (kernel_size - 1) if causal else dilation * (kernel_size - 1) // 2
    w = weight_normed_kernel(p, prefix + "/conv1d")
    xp = F.pad(x, (0, 0, pad, pad))
    out = conv1d_valid(xp, w, p[prefix + "/conv1d/bias"], dilation)
    if causal and pad != 0:
        out = out[:, :-pad]
    return out


def conv1x1(p: Params, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """Keras Conv1D kernel_size=1 (modules.py:77-97,117-127)."""
    w = weight_normed_kernel(p, prefix)
    return conv1d_valid(x, w, p[prefix + "/bias"], 1)


def conv2d_transpose_same(x: torch.Tensor, w: torch.Tensor, s: int) -> torch.Tensor:
    """[TF] conv2d_transpose, padding SAME, strides (s,1), filters 1->1 (model.py:303-309).

    x [B,H,W] (time, mel), w [2s,3].  TF defines it as the gradient of the SAME forward conv
    w.r.t. its input: out[i,m] = sum_{j,kh,kw : s*j+kh-s//2 == i, m'+kw-1 == m} x[j,m'] w[kh,kw].
    For even s this equals torch conv_transpose2d(stride=(s,1), padding=(s//2,1)).
    """
    y = F.conv_transpose2d(x[:, None], w[None, None], stride=(s, 1), padding=(s // 2, 1))
    if s % 2:  # odd stride: SAME pads (s//2, s-s//2); emulate by definition
        raise NotImplementedError("odd upsample scale")
    return y[:, 0]


def leaky_relu(x: torch.Tensor, alpha: float) -> torch.Tensor:
    """[TF] tf.nn.leaky_relu = max(x, alpha x)."""
    return torch.maximum(x, alpha * x)


# --------------------------------------------------------------------------------------
# model.py
# --------------------------------------------------------------------------------------
def upsample(p: Params, c: torch.Tensor, scales: Sequence[int]) -> torch.Tensor:
    """FloWaveNet.upsample model.py:398-404. c [B,Tm,mels] -> [B,Tm*prod(scales),mels]."""
    for n, s in enumerate(scales):
        name = "conv2d_transpose" if n == 0 else "conv2d_transpose_%d" % n
        v = p[name + "/kernel"]  # [2s,3,1,1]
        w = l2_normalize(v, (0, 2)) * p[name + "/wn/g"]  # convolutional.py:186 (per kw column)
        c = conv2d_transpose_same(c, w[:, :, 0, 0], s) + p[name + "/bias"]
        c = leaky_relu(c, 0.4)
    return c


def squeeze(x: torch.Tensor) -> torch.Tensor:
    """Block.forward model.py:226-228: out[b,t,2c+k] = x[b,2t+k,c]."""
    B, T, C = x.shape
    return x.reshape(B, T // 2, 2, C).permute(0, 1, 3, 2).reshape(B, T // 2, 2 * C)


def unsqueeze(x: torch.Tensor) -> torch.Tensor:
    """Block.reverse model.py:260-262: out[b,2t+k,c] = x[b,t,2c+k]."""
    B, T, C = x.shape
    return x.reshape(B, T, C // 2, 2).permute(0, 1, 3, 2).reshape(B, T * 2, C // 2)


def change_order(x: torch.Tensor, c: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """model.py:166-174 (g branch omitted: it has no effect on outputs)."""
    xa, xb = x.chunk(2, dim=2)
    ca, cb = c.chunk(2, dim=2)
    return torch.cat([xb, xa], 2), torch.cat([cb, ca], 2)


def actnorm_ddi(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Data-dependent init values, model.py:55-56,65-70: b = -mean(x); logs = log(1/(sqrt(mean((x+b)^2))+1e-7))/3."""
    b = -x.mean(dim=(0, 1), keepdim=True)
    var = ((x + b) ** 2).mean(dim=(0, 1), keepdim=True)
    logs = torch.log(1.0 / (torch.sqrt(var) + 1e-7)) / 3.0
    return b, logs


def actnorm_forward(p: Params, prefix: str, x: torch.Tensor, logscale: float = 3.0):
    """ActNorm.forward model.py:86-94: y=(x+b)*exp(3 logs), dlogdet=mean_c(3 logs)."""
    b = p[prefix + "/b"].to(x.dtype)
    logs = p[prefix + "/logs"].to(x.dtype) * logscale
    return (x + b) * torch.exp(logs), logs.mean()


def actnorm_reverse(p: Params, prefix: str, y: torch.Tensor, logscale: float = 3.0) -> torch.Tensor:
    """ActNorm.reverse model.py:97-102: x = y*exp(-3 logs) - b."""
    b = p[prefix + "/b"].to(y.dtype)
    logs = p[prefix + "/logs"].to(y.dtype) * logscale
    return y * torch.exp(-logs) - b


def zero_conv(p: Params, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """ZeroConv1d.forward modules.py:51-56 (no weight norm)."""
    out = conv1d_valid(x, p[prefix + "/conv1d/kernel"], p[prefix + "/conv1d/bias"], 1)
    return out * torch.exp(p[prefix + "/scale"].to(x.dtype) * 3.0)


def resblock(p: Params, prefix: str, h: torch.Tensor, c: torch.Tensor, dilation: int, causal: bool):
    """ResBlock.forward modules.py:110-128; cond conv names follow first-call order."""
    hf = conv(p, prefix + "/Conv_filter", h, 3, dilation, causal) + conv1x1(p, prefix + "/conv1d", c)
    hg = conv(p, prefix + "/Conv_gate", h, 3, dilation, causal) + conv1x1(p, prefix + "/conv1d_1", c)
    out = torch.tanh(hf) * torch.sigmoid(hg)
    res = conv1x1(p, prefix + "/conv1d_2", out)
    skip = conv1x1(p, prefix + "/conv1d_3", out)
    return (h + res) * math.sqrt(0.5), skip


def wavenet(p: Params, prefix: str, x: torch.Tensor, c: torch.Tensor, n_layer: int, causal: bool = False) -> torch.Tensor:
    """WaveNet.forward modules.py:161-186 (num_blocks=1, dilation 3**n)."""
    h = torch.relu(conv(p, prefix + "/Conv_front", x, 3, 1, causal))
    skips = []
    for n in range(n_layer):
        h, s = resblock(p, prefix + "/ResBlock_0_%d" % n, h, c, 3 ** n, causal)
        skips.append(s)
    out = torch.relu(sum(skips[1:], skips[0]))
    out = torch.relu(conv(p, prefix + "/Conv_final", out, 1, 1, causal))
    return zero_conv(p, prefix + "/ZeroConv1d", out)


def coupling_forward(p: Params, prefix: str, x, c, n_layer: int, affine: bool = True, causal: bool = False):
    """AffineCoupling.forward model.py:121-141."""
    in_a, in_b = x.chunk(2, dim=2)
    c_a, _ = c.chunk(2, dim=2)
    net = wavenet(p, prefix + "/WaveNet", in_a, c_a, n_layer, causal)
    if affine:
        log_s, t = net.chunk(2, dim=2)
        out_b = (in_b - t) * torch.exp(-log_s)
        logdet = (-log_s).mean() / 2
    else:
        out_b = in_b + net
        logdet = None
    return torch.cat([in_a, out_b], 2), logdet


def coupling_reverse(p: Params, prefix: str, y, c, n_layer: int, affine: bool = True, causal: bool = False):
    """AffineCoupling.reverse model.py:143-161."""
    out_a, out_b = y.chunk(2, dim=2)
    c_a, _ = c.chunk(2, dim=2)
    net = wavenet(p, prefix + "/WaveNet", out_a, c_a, n_layer, causal)
    if affine:
        log_s, t = net.chunk(2, dim=2)
        in_b = out_b * torch.exp(log_s) + t
    else:
        in_b = out_b - net
    return torch.cat([out_a, in_b], 2)


def flow_forward(p: Params, prefix: str, x, c, n_layer: int, affine=True, causal=False):
    """Flow.forward model.py:185-194."""
    out, logdet = actnorm_forward(p, prefix + "/ActNorm", x)
    out, det = coupling_forward(p, prefix + "/AffineCoupling", out, c, n_layer, affine, causal)
    out, c = change_order(out, c)
    if det is not None:
        logdet = logdet + det
    return out, c, logdet


def flow_reverse(p: Params, prefix: str, y, c, n_layer: int, affine=True, causal=False):
    """Flow.reverse model.py:196-202."""
    y, c = change_order(y, c)
    x = coupling_reverse(p, prefix + "/AffineCoupling", y, c, n_layer, affine, causal)
    x = actnorm_reverse(p, prefix + "/ActNorm", x)
    return x, c


def block_forward(p: Params, prefix: str, x, c, n_flow: int, n_layer: int, affine=True, causal=False):
    """Block.forward model.py:221-247."""
    out, c = squeeze(x), squeeze(c)
    logdet = 0.0
    for j in range(n_flow):
        out, c, det = flow_forward(p, "%s/Flow_%d" % (prefix, j), out, c, n_layer, affine, causal)
        logdet = logdet + det
    return out, c, logdet


def block_reverse(p: Params, prefix: str, y, c, n_flow: int, n_layer: int, affine=True, causal=False):
    """Block.reverse model.py:249-277."""
    x = y
    for j in reversed(range(n_flow)):
        x, c = flow_reverse(p, "%s/Flow_%d" % (prefix, j), x, c, n_layer, affine, causal)
    return unsqueeze(x), unsqueeze(c)


def forward(p: Params, hp, x: torch.Tensor, c: torch.Tensor, dtype=torch.float64):
    """FloWaveNet.forward model.py:317-347.  x [B,T,1], c [B,T/hop,mels].

    Returns (log_p, logdet, z_flat) where z_flat = unsqueeze^n(out) is the [B,T,1] latent
    (the reference only returns the two scalars)."""
    x = x.to(dtype)
    c = upsample({k: v.to(dtype) for k, v in p.items() if k.startswith("conv2d_transpose")}, c.to(dtype), hp.upsample_scales)
    pp = {k: v.to(dtype) for k, v in p.items()}
    out, logdet = x, 0.0
    for i in range(hp.n_block):
        out, c, det = block_forward(pp, "Block_%d" % i, out, c, hp.n_flow, hp.n_layer, hp.affine, hp.causality)
        logdet = logdet + det
    log_p = (0.5 * (-LOG_2PI - out ** 2)).mean()
    z = out
    for _ in range(hp.n_block):
        z = unsqueeze(z)
    return log_p, logdet, z


def reverse(p: Params, hp, z: torch.Tensor, c: torch.Tensor, dtype=torch.float64) -> torch.Tensor:
    """FloWaveNet.reverse model.py:350-396.  z [B,T,1] -> x [B,T,1]."""
    pp = {k: v.to(dtype) for k, v in p.items()}
    x = z.to(dtype)
    c = upsample(pp, c.to(dtype), hp.upsample_scales)
    for _ in range(hp.n_block):
        x, c = squeeze(x), squeeze(c)
    for i in reversed(range(hp.n_block)):
        x, c = block_reverse(pp, "Block_%d" % i, x, c, hp.n_flow, hp.n_layer, hp.affine, hp.causality)
    return x


def ddi_init(p: Params, hp, x: torch.Tensor, c: torch.Tensor, dtype=torch.float64) -> Params:
    """ActNorm data-dependent init pass (train.py:221,229 feed init=True; model.py:30-41).

    Runs the forward graph once, assigning every ActNorm's (b, logs) from the statistics of
    its own input, in graph order.  Returns a new params dict."""
    pp = {k: v.to(dtype).clone() for k, v in p.items()}
    out = x.to(dtype)
    cc = upsample(pp, c.to(dtype), hp.upsample_scales)
    for i in range(hp.n_block):
        out, cc = squeeze(out), squeeze(cc)
        for j in range(hp.n_flow):
            pre = "Block_%d/Flow_%d" % (i, j)
            b, logs = actnorm_ddi(out)
            pp[pre + "/ActNorm/b"], pp[pre + "/ActNorm/logs"] = b, logs
            out, cc, _ = flow_forward(pp, pre, out, cc, hp.n_layer, hp.affine, hp.causality)
    return pp


# --------------------------------------------------------------------------------------
# Parameter schema + seeded synthetic weights
# --------------------------------------------------------------------------------------
class HP:
    """Plain stand-in for the reference's tf.contrib HParams (hparams.py:6-50), path fields only."""

    def __init__(self, n_block=8, n_flow=6, n_layer=2, num_mels=80, affine=True, causality=False,
                 upsample_scales=(16, 16), gin_channels=-1, n_speakers=7, filter_size=256):
        self.n_block, self.n_flow, self.n_layer, self.num_mels = n_block, n_flow, n_layer, num_mels
        self.affine, self.causality = affine, causality
        self.upsample_scales = list(upsample_scales)
        self.gin_channels, self.n_speakers = gin_channels, n_speakers
        self.filter_size = filter_size

    @property
    def hop(self):
        h = 1
        for s in self.upsample_scales:
            h *= s
        return h


def param_shapes(hp) -> Dict[str, Tuple[int, ...]]:
    """Every variable the reference creates for the path, with its shape (names: SURVEY 8f-3)."""
    F_ = hp.filter_size
    shapes: Dict[str, Tuple[int, ...]] = {}
    for n, s in enumerate(hp.upsample_scales):
        name = "conv2d_transpose" if n == 0 else "conv2d_transpose_%d" % n
        shapes[name + "/kernel"] = (2 * s, 3, 1, 1)
        shapes[name + "/wn/g"] = (1,)
        shapes[name + "/bias"] = (1,)
    cx, cc = 1, hp.num_mels
    for i in range(hp.n_block):
        cx, cc = cx * 2, cc * 2
        out_ch = cx if hp.affine else cx // 2
        for j in range(hp.n_flow):
            pre = "Block_%d/Flow_%d" % (i, j)
            shapes[pre + "/ActNorm/b"] = (1, 1, cx)
            shapes[pre + "/ActNorm/logs"] = (1, 1, cx)
            w = pre + "/AffineCoupling/WaveNet"

            def convp(name, k, cin, cout):
                shapes[name + "/kernel"] = (k, cin, cout)
                shapes[name + "/wn/g"] = (cout,)
                shapes[name + "/bias"] = (cout,)

            convp(w + "/Conv_front/conv1d", 3, cx // 2, F_)
            for n in range(hp.n_layer):
                r = w + "/ResBlock_0_%d" % n
                convp(r + "/Conv_filter/conv1d", 3, F_, F_)
                convp(r + "/Conv_gate/conv1d", 3, F_, F_)
                convp(r + "/conv1d", 1, cc // 2, F_)
                convp(r + "/conv1d_1", 1, cc // 2, F_)
                convp(r + "/conv1d_2", 1, F_, F_)
                convp(r + "/conv1d_3", 1, F_, F_)
            convp(w + "/Conv_final/conv1d", 1, F_, F_)
            shapes[w + "/ZeroConv1d/conv1d/kernel"] = (1, F_, out_ch)
            shapes[w + "/ZeroConv1d/conv1d/bias"] = (out_ch,)
            shapes[w + "/ZeroConv1d/scale"] = (1, 1, out_ch)
    return shapes


def synthetic_params(hp, seed: int = 0, dtype=torch.float32) -> Params:
    """Seeded weights mirroring the reference initialisers (he-uniform kernels and biases,
    wn/g ~ U(0.5,1.5) to exercise weight-norm, small NON-zero ZeroConv so couplings are
    non-trivial; ActNorm ~ small random, overwritten by ddi_init where a test needs it)."""
    import numpy as np

    rng = np.random.default_rng(seed)
    out: Params = {}
    for name, shp in param_shapes(hp).items():
        if name.endswith("/kernel"):
            if "ZeroConv1d" in name:
                a = rng.uniform(-0.02, 0.02, shp)
            else:
                fan_in = int(np.prod(shp[:-1])) if len(shp) == 3 else shp[0] * shp[1]
                lim = math.sqrt(6.0 / fan_in)
                a = rng.uniform(-lim, lim, shp)
        elif name.endswith("/wn/g"):
            a = rng.uniform(0.5, 1.5, shp)
        elif name.endswith("/bias"):
            if "ZeroConv1d" in name:
                a = rng.uniform(-0.02, 0.02, shp)
            elif name.startswith("conv2d_transpose"):
                a = rng.uniform(-0.05, 0.05, shp)
            else:
                a = rng.uniform(-1, 1, shp) * math.sqrt(6.0 / shp[0]) * 0.1
        elif name.endswith("/scale"):
            a = rng.uniform(-0.1, 0.1, shp)
        elif name.endswith("/ActNorm/b"):
            a = rng.uniform(-0.1, 0.1, shp)
        elif name.endswith("/ActNorm/logs"):
            a = rng.uniform(-0.05, 0.05, shp)
        else:
            raise KeyError(name)
        out[name] = torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
    return out


def synthetic_inputs(hp, B: int, n_frames: int, seed: int, kind: str):
    """SURVEY 8d inputs: mel ~ U[0,1); kind='z': N(0,1)*0.7; kind='x': chirp + noise, clipped."""
    import numpy as np

    rng = np.random.default_rng(seed)
    T = n_frames * hp.hop
    c = rng.uniform(0.0, 1.0, (B, n_frames, hp.num_mels)).astype(np.float32)
    if kind == "z":
        a = (rng.standard_normal((B, T, 1)) * 0.7).astype(np.float32)
    else:
        t = np.arange(T, dtype=np.float64)[None, :, None] / T
        ph = rng.uniform(0, 2 * np.pi, (B, 1, 1))
        a = 0.5 * np.sin(2 * np.pi * (200.0 * t + 800.0 * t * t) + ph) + 0.1 * rng.standard_normal((B, T, 1))
        a = np.clip(a, -0.999, 0.999).astype(np.float32)
    return torch.from_numpy(a), torch.from_numpy(c)
