#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
Q="--steps 30 --warmup 5 --no-latency --no-cpu-baseline --no-widen"
run() {  # name, env..., args
  name=$1; shift
  env "$@" timeout 600 python bench.py $Q $EXTRA > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_$name.json").read().strip().splitlines()[-1])
    print("$name: value", d["value"], "ms/step", d["ms_per_step"], "e2e", d.get("e2e", {}).get("value"))
except Exception as e:
    print("$name failed", e, open("gpurun_out/ab_$name.err").read()[-400:])
PY
}
EXTRA="" run async0_r2 SVOB200_TRACKER_ASYNC_DF=0 SVOB200_TRACKER_RANGES=2
EXTRA="" run async1_r1 SVOB200_TRACKER_ASYNC_DF=1 SVOB200_TRACKER_RANGES=1
EXTRA="" run async1_r2 SVOB200_TRACKER_ASYNC_DF=1 SVOB200_TRACKER_RANGES=2
EXTRA="" run async1_r4 SVOB200_TRACKER_ASYNC_DF=1 SVOB200_TRACKER_RANGES=4
EXTRA="--seqs 512 --no-e2e" run s512_async0 SVOB200_TRACKER_ASYNC_DF=0 SVOB200_TRACKER_RANGES=2
EXTRA="--seqs 512 --no-e2e" run s512_async1_r1 SVOB200_TRACKER_ASYNC_DF=1 SVOB200_TRACKER_RANGES=1
EXTRA="--seqs 512 --no-e2e" run s512_async1_r2 SVOB200_TRACKER_ASYNC_DF=1 SVOB200_TRACKER_RANGES=2
EXTRA="--seqs 1024 --no-e2e" run s1024_async1_r2 SVOB200_TRACKER_ASYNC_DF=1 SVOB200_TRACKER_RANGES=2
EXTRA="--seqs 2048 --no-e2e" run s2048_async1_r2 SVOB200_TRACKER_ASYNC_DF=1 SVOB200_TRACKER_RANGES=2
EXTRA="--seed-regime young --no-e2e" run young SVOB200_TRACKER_ASYNC_DF=1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/ab_young.json").read().strip().splitlines()[-1])
print("young:", d["seed_workload"], {k: v["ms"] for k, v in d["stages"].items() if v["ms"] > 0.1}, d.get("zmssd_evals_per_s"))
d = json.loads(open("gpurun_out/ab_async1_r2.json").read().strip().splitlines()[-1])
print("steady:", d["seed_workload"], d.get("zmssd_evals_per_s"), d["clocks"])
PY
