#!/bin/bash
# usage (GPU box): bash tools/layer_ab.sh "<env assignments>" ...   -- C3 pass time and the fused-layer family time for each setting
mkdir -p gpurun_out
for cfg in "$@"; do
  env $cfg timeout 200 python bench.py --workload ${WL:-c3} --steps ${STEPS:-6} --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - "$cfg" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/ab.json")); f = d["roofline"]["families"]
    print("%-40s pass %.2f ms  %s  sm %s MHz" % (sys.argv[1], d["ms_per_step"], "  ".join("%s %.2f" % (k.split("_")[0], v["ms"] / d["steps"]) for k, v in f.items()), d["clocks"]["sm_mhz"]))
except Exception as e:
    print(sys.argv[1], "ERR", e, open("gpurun_out/ab.err").read()[-500:])
PY
done
