#!/bin/bash
# A/B of library variants: bash tools/r2_ab_lib.sh <variant> ... (android_svo_b200/lib/libsvob200_<variant>.so; "base" = the in-tree build)
mkdir -p gpurun_out
for rep in 1 2; do
for lib in "$@"; do
  L=/root/repo/android_svo_b200/lib/libsvob200_$lib.so; [ $lib = base ] && L=/root/repo/android_svo_b200/lib/libsvob200.so
  for seqs in 4096 512; do
    SVOB200_LIB=$L timeout 600 python bench.py --seqs $seqs --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-widen --no-latency > gpurun_out/ab_${lib}_$seqs.json 2> gpurun_out/ab_${lib}_$seqs.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/ab_${lib}_$seqs.json").read().strip().splitlines()[-1])
print("$lib seqs $seqs value %.0f ms/step %.4f" % (d["value"], d["ms_per_step"]), {k: v["ms"] for k, v in d["stages"].items()})
PY
  done
done
done
