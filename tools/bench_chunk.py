"""Reverse pass of one C4 time chunk as a rank of an N-GPU job sees it (hparams.py model, one utterance of T/N + 2 halos samples):
   python tools/bench_chunk.py [n_ranks=8]      -- ms per pass and launches, for A/B runs of launch-level heuristics (FWN_* env)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import tf_flowavenet_b200 as P
from tf_flowavenet_b200.synthetic import synthetic_inputs, synthetic_params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
hp = P.HParams(**{**P.hparams.values(), "dtype": "bfloat16"})
net = P.FloWaveNet(hp, variables=P.VariableStore())
net.load_variables(synthetic_params(net.variable_shapes(), seed=1234))
hop, halo = 256, 15616
frames = (1323008 // n + 2 * halo) // hop
z, c = synthetic_inputs(hop, 80, 1, frames, 7, "z")
z, c = torch.from_numpy(z).cuda(), torch.from_numpy(c).cuda()
for _ in range(4):
    net.reverse(z, c)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    net.reverse(z, c)
e1.record()
torch.cuda.synchronize()
print("chunk of a %d-rank job: T=%d  %.3f ms per pass  %d launches  [%s]" % (n, z.shape[1], e0.elapsed_time(e1) / reps, net.last_launches(),
      " ".join("%s=%s" % (k, v) for k, v in os.environ.items() if k.startswith("FWN_"))))
