#!/bin/bash
# round-2 evidence run on the GPU box: tests, smoke, both bench arms, the extra regimes.  TAG=$1
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -n "passed\|failed\|keyframe insertion" gpurun_out/${TAG}_pytest_gpu.log | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_reference_n1.json 2> gpurun_out/${TAG}_reference_n1.err; echo "reference rc=$?"
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"
Q="--steps 20 --warmup 5 --no-latency --no-widen"
timeout 900 python bench.py $Q --seed-regime young > gpurun_out/${TAG}_young_bench_n1.json 2> gpurun_out/${TAG}_young.err; echo "young rc=$?"
timeout 900 python bench.py $Q --keyframe-every 5 > gpurun_out/${TAG}_keyframes_bench_n1.json 2> gpurun_out/${TAG}_keyframes.err; echo "keyframes rc=$?"
timeout 900 python bench.py $Q --chain > gpurun_out/${TAG}_chain_bench_n1.json 2> gpurun_out/${TAG}_chain.err; echo "chain rc=$?"
timeout 600 python bench.py --impl dropin --steps 6 > gpurun_out/${TAG}_dropin.json 2> gpurun_out/${TAG}_dropin.err; echo "dropin rc=$?"
timeout 600 python bench.py --impl dropin --steps 6 --chain > gpurun_out/${TAG}_dropin_chain.json 2> gpurun_out/${TAG}_dropin_chain.err; echo "dropin chain rc=$?"
python - <<PY
import json
def last(f):
    try: return json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: return {"error": repr(e)}
d = last("gpurun_out/${TAG}_bench_n1.json")
if "value" in d:
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "cpu", d.get("cpu_baseline", {}).get("value"), d["clocks"])
    print({k: v["ms"] for k, v in d["stages"].items()})
    print({k: (v.get("ms") if isinstance(v, dict) else v) for k, v in d.get("next_rows", {}).items()}, d.get("next_rows", {}).get("fast_detect", {}).get("natural_texture"))
    for n, v in d.get("latency", {}).items(): print(n, v.get("resident"), v.get("host_buffers"), v.get("cpu_reference_1_thread"))
    print(d["roofline"]["frac"], d["roofline_pyramid"]["frac"], d["seed_workload"])
else: print(d, open("gpurun_out/${TAG}_bench_n1.err").read()[-800:])
for n in ("young", "keyframes", "chain"):
    d = last("gpurun_out/${TAG}_%s_bench_n1.json" % n)
    print(n, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"), d.get("seed_workload", {}).get("n_evals_mean"), d.get("error"))
print("reference", last("gpurun_out/${TAG}_reference_n1.json").get("value"))
d = last("gpurun_out/${TAG}_dropin.json"); print("dropin", d.get("dropin"), d.get("reference"), d.get("speedup_p50"), d.get("error"))
d = last("gpurun_out/${TAG}_dropin_chain.json"); print("dropin chain", d.get("dropin"), d.get("reference"), d.get("speedup_p50"), d.get("error"))
print("keyframes", last("gpurun_out/${TAG}_keyframes_bench_n1.json").get("keyframes"))
PY
