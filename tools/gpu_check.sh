#!/bin/bash
# GPU-box check: the -m gpu suite (all failures, with the parity counters printed), then one bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
if [ -n "$SANITIZE" ]; then
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer.log 2>&1
  echo "sanitizer rc=$?"; tail -25 gpurun_out/sanitizer.log
fi
timeout 1500 python -m pytest tests -q -m gpu -s -rA --durations=15 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d.get("e2e", {}).get("value"))
    print({k: v["ms"] for k, v in d["stages"].items()})
    print({k: (v.get("ms") if isinstance(v, dict) else v) for k, v in d.get("next_rows", {}).items()})
    print({k: (v.get("resident", {}).get("p50_ms"), v.get("host_buffers", {}).get("p50_ms")) for k, v in d.get("latency", {}).items() if isinstance(v, dict)})
except Exception as e:
    print("bench parse failed", e)
PY
