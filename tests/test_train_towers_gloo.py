"""Data-parallel towers on CPU (gloo, world_size 2): the tower average of the training step (utils.py:34-60) as the product
implements it -- one all-reduce of the flat gradient vector -- against the oracle's per-variable average, followed by the
oracle's clip + Adam so that every rank ends on identical variables (train.py:70-81)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import flowavenet_oracle as O
from oracle import flowavenet_train_oracle as TO
from tf_flowavenet_b200.train import (average_flat_gradients, average_flat_gradients_bucketed, broadcast_flat_variables, bucket_ranges,
                                      learning_rate)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        hp = O.HP(n_block=2, n_flow=2, n_layer=1, num_mels=4, upsample_scales=(2, 2))
        params = O.synthetic_params(hp, 3, torch.float64)          # same variables on every tower
        batches = [O.synthetic_inputs(hp, 1, 4, 50 + r, "x") for r in range(world)]   # each tower draws its own batch
        _, _, _, mine = TO.loss_and_grads(params, hp, *batches[rank])
        names = sorted(mine)
        flat = torch.cat([mine[k].reshape(-1) for k in names])
        average_flat_gradients(flat)                                # the product's tower average
        ref = TO.average_gradients([TO.loss_and_grads(params, hp, *b)[3] for b in batches])
        ref_flat = torch.cat([ref[k].reshape(-1) for k in names])
        err = float((flat - ref_flat).abs().max())
        # the bucketed average (what the GPU trainer enqueues per block while the backward pass is still running) is the same
        # reduction cut at the bucket borders: buckets tile the flat vector exactly and each is averaged on its own
        shapes = O.param_shapes(hp)
        buckets = bucket_ranges(shapes, hp.n_block)
        total = sum((int(torch.tensor(s).prod()) + 3) & ~3 for s in shapes.values())
        cover = sorted(buckets)
        tiled = cover[0][0] == 0 and all(a[0] + a[1] == b[0] for a, b in zip(cover, cover[1:])) and cover[-1][0] + cover[-1][1] == total
        padded = torch.zeros(total, dtype=torch.float64)
        off = 0
        for k, s in shapes.items():      # the library's flat layout: variables in schema order, each padded to 4 floats
            n = mine[k].numel()
            padded[off:off + n] = mine[k].reshape(-1)
            off += (n + 3) & ~3
        one_shot = padded.clone()
        average_flat_gradients(one_shot)
        average_flat_gradients_bucketed(padded, buckets)
        err = max(err, float((padded - one_shot).abs().max()), 0.0 if tiled else 1.0)
        err = max(err, 0.0 if buckets[0][0] == max(b[0] for b in buckets[:hp.n_block]) else 1.0)   # last block is produced first
        # clip + Adam on the averaged gradient: all ranks must agree bit for bit
        clipped, norm = TO.clip_by_global_norm({"g": flat}, 1.0)
        m, v = {"g": torch.zeros_like(flat)}, {"g": torch.zeros_like(flat)}
        new = TO.adam_step({"g": torch.zeros_like(flat)}, clipped, m, v, learning_rate(0), 1)["g"]
        gathered = [torch.zeros_like(new) for _ in range(world)]
        dist.all_gather(gathered, new)
        same = all(torch.equal(gathered[0], t) for t in gathered)
        # after the data-dependent init every tower adopts rank 0's variables
        mine_vars = torch.full((7,), float(rank + 1))
        broadcast_flat_variables(mine_vars)
        same = same and bool(torch.all(mine_vars == 1.0))
        q.put((rank, err, same, norm))
    finally:
        dist.destroy_process_group()


def test_tower_average_matches_oracle():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, same, norm in res:
        assert err < 1e-15, (rank, err)
        assert same
        assert norm > 0


def test_learning_rate_schedule_matches_oracle():
    for step in (0, 1, 199999, 200000, 399999, 400000, 599999, 600000, 10 ** 6):
        assert learning_rate(step) == TO.learning_rate(step)
