#!/bin/bash
# round-2 A/B on the GPU box: tests, range pipelining on/off, ncu of the FAST kernel and the seed kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
Q="--steps 30 --warmup 5 --no-latency --no-cpu-baseline --no-widen"
for R in 1 2 4; do
  SVOB200_TRACKER_RANGES=$R timeout 600 python bench.py $Q > gpurun_out/ab_ranges$R.json 2> gpurun_out/ab_ranges$R.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/ab_ranges$R.json").read().strip().splitlines()[-1])
print("ranges $R: value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], {k: v["ms"] for k, v in d["stages"].items() if v["ms"] > 0.1})
PY
done
for S in 512 1024; do
for R in 1 2; do
  SVOB200_TRACKER_RANGES=$R timeout 600 python bench.py $Q --seqs $S --no-e2e > gpurun_out/ab_s${S}_r$R.json 2> gpurun_out/ab_s${S}_r$R.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/ab_s${S}_r$R.json").read().strip().splitlines()[-1])
print("seqs $S ranges $R: value", d["value"], "ms/step", d["ms_per_step"])
PY
done
done
python tools/fast_probe.py 1024 > gpurun_out/fast_probe.txt 2>&1; cat gpurun_out/fast_probe.txt
ncu --set full --clock-control none --import-source on -k regex:fast_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r2a_fast python tools/fast_probe.py 256 > gpurun_out/ncu_fast.log 2>&1
ncu -i gpurun_out/r2a_fast.ncu-rep --page raw --csv > gpurun_out/r2a_fast_raw.csv
ncu -i gpurun_out/r2a_fast.ncu-rep --page source --csv > gpurun_out/r2a_fast_source.csv 2>/dev/null
CMD="python bench.py --seqs 4096 --steps 2 --warmup 1 --no-latency --no-cpu-baseline --no-e2e --no-widen"
SVOB200_TRACKER_RANGES=1 ncu --set full --clock-control none --import-source on -k regex:'seeds_geom|seeds_finish' --launch-skip 4 -c 2 -f -o gpurun_out/r2a_seeds $CMD > gpurun_out/ncu_seeds.log 2>&1
ncu -i gpurun_out/r2a_seeds.ncu-rep --page raw --csv > gpurun_out/r2a_seeds_raw.csv
ls -la gpurun_out | tail -20
