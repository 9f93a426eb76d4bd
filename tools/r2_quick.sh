#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "keyframe or device_mode or chain or golden" > gpurun_out/quick_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/quick_pytest.log
for i in 1 2; do
timeout 600 python bench.py --steps 30 --warmup 5 --no-latency --no-cpu-baseline --no-widen --no-e2e > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/quick_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], {k: v["ms"] for k, v in d["stages"].items()}, d["roofline_pyramid"]["frac"])
PY
done
