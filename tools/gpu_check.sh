#!/bin/bash
# Staged GPU check with tight per-stage timeouts: a hung kernel costs one stage's timeout, not the whole budget.
# usage (on the GPU box): bash tools/gpu_check.sh [stage ...]   stages: mixed train16 dbg16 baseline train model bench
mkdir -p gpurun_out
run() {  # name timeout cmd...
  local name=$1 t=$2; shift 2
  timeout "$t" "$@" > gpurun_out/chk_$name.log 2>&1
  local rc=$?
  echo "== $name rc=$rc"
  grep -E "^\.?(bf16|loss|wgrad|C2|C3|deep|C4|mixed|step|    Block|    conv)|passed|failed|Error|error" gpurun_out/chk_$name.log | tail -${TAILN:-25} | cut -c1-330
  if [ $rc -eq 124 ]; then echo "!! $name TIMED OUT -- stopping"; exit 0; fi
}
for s in "$@"; do
  case $s in
    mixed) run mixed 240 python -m pytest tests/test_gpu_mixed.py -m gpu -q -s -x ;;
    dbg16) FWN_TRAIN_GRAPH=0 FWN_TRACE=1 FWN_SYNC_DEBUG=1 run dbg16 100 python tools/debug_train16.py 8 8 25; tail -2 gpurun_out/chk_dbg16.log | cut -c1-300 ;;
    train16) run train16 400 python -m pytest tests/test_gpu_train_bf16.py -m gpu -q -s ;;
    baseline) run baseline 500 python -m pytest tests/test_gpu_baseline_shapes.py -m gpu -q -s ;;
    train) run train 500 python -m pytest tests/test_gpu_train.py -m gpu -q -s ;;
    model) run model 400 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -m gpu -q -s ;;
    bench_*) w=${s#bench_}; timeout 300 python bench.py --workload ${w%%,*} --steps 10 --warmup 3 --no-cpu-baseline $BENCH_ARGS > gpurun_out/chk_$s.json 2> gpurun_out/chk_$s.err
             echo "== $s rc=$?"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/chk_$s.json")); r = d["roofline"]
    print("   ms/step %.3f  value %.4g  frac %.3f  launches %s  e2e %.4g" % (d["ms_per_step"], d["value"], r["frac"], d.get("gpu_launches"), d["e2e"]["value"]))
    if "phases_ms_per_step" in r: print("   phases", r["phases_ms_per_step"])
    if "families" in r: print("   " + "  ".join("%s %.2f ms" % (k, v["ms"] / d["steps"]) for k, v in r["families"].items()))
except Exception as e:
    print("   ERR", e); print(open("gpurun_out/chk_$s.err").read()[-600:])
PY
             ;;
    *) echo "unknown stage $s" ;;
  esac
done
