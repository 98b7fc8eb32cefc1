"""Time-chunk sharding of one long utterance across ranks with receptive-field halos (BASELINE config C4).

Every rank owns a contiguous chunk of z [B, T/N, 1] and of the mel frames [B, T/(N*hop), mels].  Before the pass it
exchanges `halo` samples of z (and halo/hop mel frames) with rank-1 and rank+1 (point-to-point: NCCL send/recv over
NVLink on GPUs, gloo on CPU), runs the inverse pass on the extended chunk and keeps the interior
(`fwn_reverse_chunk`).  With halo >= receptive_halo(hparams) the result is bit-for-bit the pass a single device would
have produced for those samples: chunk-interior halos carry real neighbour data while true utterance edges keep the
reference's zero padding (modules.py:27).  No collective touches the data path.
"""
import math


def receptive_halo(hp):
    """Samples of context per side that the inverse (or forward) pass of one output sample can see.

    Per coupling WaveNet the receptive half-width is 1 (front conv) + sum_n 3^n (layers) squeezed steps (twice that for
    causal nets); a step of block i spans 2^(i+1) samples; plus one hop for the two transposed convs of the upsampler.
    Rounded up to a whole number of mel frames and squeeze groups.  Mirrors fwn_receptive_halo()."""
    rw = 1 + sum(3 ** n for n in range(hp.n_layer))
    if getattr(hp, "causality", False):
        rw *= 2
    hop = 1
    for s in hp.upsample_scales:
        hop *= s
    halo = sum(hp.n_flow * rw * (2 << i) for i in range(hp.n_block)) + hop
    q = hop * (1 << hp.n_block) // math.gcd(hop, 1 << hp.n_block)
    return (halo + q - 1) // q * q


def chunk_bounds(T, n_ranks, quantum):
    """Interior [lo, hi) of every rank: equal chunks, multiples of `quantum` (= lcm(hop, 2^n_block))."""
    if T % (n_ranks * quantum):
        raise ValueError("T=%d must be a multiple of n_ranks*quantum=%d" % (T, n_ranks * quantum))
    step = T // n_ranks
    return [(r * step, (r + 1) * step) for r in range(n_ranks)]


def exchange_halos(local, halo, rank, world, group=None):
    """local: [B, L, C] tensor (any device).  Returns (left, right): `halo` rows received from rank-1 / rank+1, or None at
    the utterance edges.  Uses batched point-to-point ops (ncclSend/ncclRecv under NCCL, gloo send/recv on CPU)."""
    import torch
    import torch.distributed as dist
    if halo > local.shape[1]:
        raise ValueError("halo %d exceeds the local chunk length %d" % (halo, local.shape[1]))
    ops, left, right = [], None, None
    if rank > 0:
        left = torch.empty_like(local[:, :halo])
        send_l = local[:, :halo].contiguous()
        ops += [dist.P2POp(dist.isend, send_l, rank - 1, group), dist.P2POp(dist.irecv, left, rank - 1, group)]
    if rank < world - 1:
        right = torch.empty_like(local[:, :halo])
        send_r = local[:, -halo:].contiguous()
        ops += [dist.P2POp(dist.isend, send_r, rank + 1, group), dist.P2POp(dist.irecv, right, rank + 1, group)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return left, right


def extend_with_halos(local, left, right):
    import torch
    parts = [p for p in (left, local, right) if p is not None]
    return torch.cat(parts, dim=1).contiguous(), (0 if left is None else left.shape[1]), (0 if right is None else right.shape[1])


def reverse_sharded(run_chunk, hp, z_local, c_local, rank, world, group=None, halo=None):
    """Inverse pass of this rank's time chunk.  `run_chunk(z_ext, c_ext, halo_l, halo_r) -> x_interior` is the device
    routine (FloWaveNet.reverse_chunk).  Returns x for the local interior [B, T/N, 1]."""
    hop = 1
    for s in hp.upsample_scales:
        hop *= s
    halo = receptive_halo(hp) if halo is None else halo
    if world == 1:
        return run_chunk(z_local, c_local, 0, 0)
    zl, zr = exchange_halos(z_local, halo, rank, world, group)
    cl, cr = exchange_halos(c_local, halo // hop, rank, world, group)
    z_ext, hl, hr = extend_with_halos(z_local, zl, zr)
    c_ext, _, _ = extend_with_halos(c_local, cl, cr)
    return run_chunk(z_ext, c_ext, hl, hr)
