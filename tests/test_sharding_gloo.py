"""Multi-process (gloo, world_size 2 and 3, CPU) tests of the time-chunk sharding host logic: halo exchange over
point-to-point ops and exactness of overlap-recompute chunking, with the oracle standing in for the device routine."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import flowavenet_oracle as O
from tf_flowavenet_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, frames_per_rank, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        hp = O.HP(n_block=2, n_flow=2, n_layer=2, num_mels=4, upsample_scales=(2, 2))
        params = O.synthetic_params(hp, 3, torch.float64)
        z, c = O.synthetic_inputs(hp, 2, frames_per_rank * world, 4, "z")  # every rank derives the same full input
        T, hop = z.shape[1], hp.hop
        lo, hi = sharding.chunk_bounds(T, world, 4)[rank]
        z_loc, c_loc = z[:, lo:hi].contiguous(), c[:, lo // hop:hi // hop].contiguous()
        halo = sharding.receptive_halo(hp)
        # 1) the exchange delivers exactly the neighbours' boundary samples
        zl, zr = sharding.exchange_halos(z_loc, halo, rank, world)
        assert (zl is None) == (rank == 0) and (zr is None) == (rank == world - 1)
        if zl is not None:
            assert torch.equal(zl, z[:, lo - halo:lo])
        if zr is not None:
            assert torch.equal(zr, z[:, hi:hi + halo])

        # 2) overlap-recompute with that halo reproduces the un-sharded pass on the interior
        def run_chunk(z_ext, c_ext, hl, hr):
            x = O.reverse(params, hp, z_ext, c_ext, torch.float64)
            return x[:, hl:x.shape[1] - hr]

        x_loc = sharding.reverse_sharded(run_chunk, hp, z_loc.double(), c_loc.double(), rank, world)
        full = O.reverse(params, hp, z, c, torch.float64)[:, lo:hi]
        err = float((x_loc - full).abs().max())
        # 3) and a halo that is one quantum too short does NOT (the receptive field is tight, not padded)
        short = sharding.reverse_sharded(run_chunk, hp, z_loc.double(), c_loc.double(), rank, world, halo=halo - 16)
        err_short = float((short - full).abs().max())
        q.put((rank, err, err_short, halo))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_reverse_matches_unsharded(world):
    hp = O.HP(n_block=2, n_flow=2, n_layer=2, num_mels=4, upsample_scales=(2, 2))
    halo = sharding.receptive_halo(hp)
    assert halo == 2 * 5 * (2 + 4) + 4  # sum_blocks n_flow*rw*2^(i+1) + hop, already a multiple of lcm(hop, 2^n_block)=4
    frames_per_rank = (halo + 16) // 4 + 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, frames_per_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, err_short, _ in res:
        assert err < 1e-12, (rank, err)
    assert max(e for _, _, e, _ in res) > 1e-9  # the too-short halo is visibly wrong on at least one rank


def test_halo_formula_matches_survey():
    assert sharding.receptive_halo(O.HP()) == 15616                     # 15300 + hop 256 -> 61 frames
    assert sharding.receptive_halo(O.HP(n_block=5, upsample_scales=(8, 12))) == 2016  # 1860 + 96 -> 21 frames
    with pytest.raises(ValueError):
        sharding.chunk_bounds(1000, 3, 256)
