"""One C3-shaped FORWARD pass (block 0 first) for ncu: warm-up eagerly, then profile the first kernels of the second pass.

  FWN_GRAPH=0 ncu --set full --clock-control none --import-source on --profile-from-start off -c 12 -o gpurun_out/r2_c3 python tools/ncu_c3.py

Kernel order after cudaProfilerStart: upsample x2, front_pack, front GEMM, gate, res|skip, gate, res|skip, final, zero/affine, ...
Numbers under ncu are never bench values."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import tf_flowavenet_b200 as P
from tf_flowavenet_b200.synthetic import synthetic_inputs, synthetic_params

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
hp = P.HParams(**{**P.hparams8000.values(), "dtype": os.environ.get("DTYPE", "bfloat16")})
net = P.FloWaveNet(hp, variables=P.VariableStore())
net.load_variables(synthetic_params(net.variable_shapes(), seed=1234))
x, c = synthetic_inputs(96, 80, B, 834, 7, "x")
x, c = torch.from_numpy(x).cuda(), torch.from_numpy(c).cuda()
net.forward(x, c)
torch.cuda.synchronize()
torch.cuda.profiler.start()
net.forward(x, c)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
