#!/usr/bin/env python
"""One steady-state training step of the C5 workload between cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python tools/prof_train.py
and a summariser: python tools/prof_train.py --summarise gpurun_out/x.csv"""
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def summarise(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("fwn::", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.2f | %.1f%% | %.1f |" % (k[:80], v[0], v[1], 100 * v[1] / tot, v[1] / v[0] * 1e3))
    print("| total | %d | %.2f | | |" % (sum(v[0] for v in agg.values()), tot))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--summarise":
        return summarise(sys.argv[2])
    import numpy as np
    import torch
    import tf_flowavenet_b200 as P
    from tf_flowavenet_b200.synthetic import synthetic_inputs, synthetic_params
    from tf_flowavenet_b200.train import Trainer
    hp = P.hparams
    hop = int(np.prod(hp.upsample_scales))
    B, n_frames = 8, 25
    net = P.FloWaveNet(P.HParams(**{**hp.values(), "dtype": "float32", "gin_channels": 16, "n_speakers": 7}), variables=P.VariableStore())
    net.load_variables(synthetic_params(net.variable_shapes(), seed=1234))
    x, c = synthetic_inputs(hop, 80, B, n_frames, 1239, "x")
    g = torch.zeros(B, dtype=torch.int32).cuda()
    x, c = torch.from_numpy(x).cuda(), torch.from_numpy(c).cuda()
    tr = Trainer(net, split_terms=3, compute_dtype=os.environ.get("TRAIN_DTYPE", "bfloat16"))
    tr.train_step(x, c, g, init=True)
    tr.train_step(x, c, g)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    tr.train_step(x, c, g)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
