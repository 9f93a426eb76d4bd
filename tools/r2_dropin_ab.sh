#!/bin/bash
# drop-in variants: one or two contexts (SVOB200_DROPIN_DF_CTX), frame mirrors rebuilt on the device or uploaded level by level
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dropin.py -q -m gpu 2>&1 | tail -2
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  for mode in "" "--chain"; do
    SVOB200_DROPIN_DF_CTX=$1 SVOB200_DROPIN_REBUILD_PYRAMID=$2 timeout 400 python bench.py --impl dropin --steps 6 $mode > gpurun_out/dropin_ab.json 2> gpurun_out/dropin_ab.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/dropin_ab.json").read().strip().splitlines()[-1])
t=d["two_threads"]
print("df_ctx=$1 rebuild=$2 mode='$mode' 1-thread p50 dropin %.4f ref %.4f (x%.2f) ops %s | 2-thread ms/frame dropin %.4f ref %.4f (x%.2f)" % (d["dropin"]["p50_ms"], d["reference"]["p50_ms"], d["speedup_p50"], d["dropin"]["per_operator_mean"], t["dropin_ms_per_frame"], t["reference_ms_per_frame"], t["speedup"]))
PY
  done
done
