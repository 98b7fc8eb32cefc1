"""What a per-flow halo exchange (BASELINE config 4, "scheme 2") would cost: 48 DEPENDENT neighbour exchanges per pass (one per flow,
a few rows of the flow variable each way) against the one exchange of input halos the library does ("scheme 1", overlap-recompute).
   torchrun --nproc-per-node 2 tools/bench_halo_latency.py
Prints the time of 48 back-to-back bidirectional NCCL send/recv pairs of 4 KB (stream-ordered, no host sync in between)."""
import os

import torch
import torch.distributed as dist

dist.init_process_group("nccl")
rank, world = dist.get_rank(), dist.get_world_size()
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
peer = rank ^ 1
send = torch.zeros(1024, device="cuda")
recv = torch.zeros(1024, device="cuda")
work = torch.zeros(1 << 20, device="cuda")


def exchange():
    ops = [dist.P2POp(dist.isend, send, peer), dist.P2POp(dist.irecv, recv, peer)]
    for r in dist.batch_isend_irecv(ops):
        r.wait()          # stream-ordered wait: the next kernel on this stream sees recv


def one_pass(n=48):
    for _ in range(n):
        exchange()
        work.add_(recv[0])   # a dependent kernel between exchanges, as a flow would be


for _ in range(5):
    one_pass()
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    one_pass()
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("48 dependent 4 KB neighbour exchanges (NCCL send/recv over NVLink, %d ranks): %.3f ms per pass = %.1f us per exchange" %
          (world, float(t), float(t) * 1e3 / 48))
dist.destroy_process_group()
