#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dropin.py tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -2
timeout 400 python bench.py --impl dropin --steps 6 > gpurun_out/r2c_dropin.json 2> gpurun_out/r2c_dropin.err; echo rc=$?
timeout 400 python bench.py --impl dropin --steps 6 --chain > gpurun_out/r2c_dropin_chain.json 2> gpurun_out/r2c_dropin_chain.err; echo rc=$?
SVOB200_DROPIN_DF_CTX=1 timeout 400 python bench.py --impl dropin --steps 6 --chain > gpurun_out/r2c_dropin_chain_2ctx.json 2> gpurun_out/r2c_dropin_chain_2ctx.err; echo rc=$?
python - <<PY
import json
for f in ("r2c_dropin","r2c_dropin_chain","r2c_dropin_chain_2ctx"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); t=d["two_threads"]
    print(f, "1-thread p50 dropin %.4f ref %.4f (x%.2f) ops %s | 2-thread ms/frame dropin %.4f ref %.4f (x%.2f)" % (d["dropin"]["p50_ms"], d["reference"]["p50_ms"], d["speedup_p50"], d["dropin"]["per_operator_mean"], t["dropin_ms_per_frame"], t["reference_ms_per_frame"], t["speedup"]))
PY
