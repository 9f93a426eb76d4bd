#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/all_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/all_pytest.log
SVOB200_LIB=/root/repo/android_svo_b200/lib/libsvob200_T.so timeout 300 python tools/align_timing.py 2>&1 | tail -3
for lib in c4 c5 c6 c4; do
  L=/root/repo/android_svo_b200/lib/libsvob200_$lib.so; [ $lib = c4 ] && L=/root/repo/android_svo_b200/lib/libsvob200.so
  for seqs in 4096 512; do
    SVOB200_LIB=$L timeout 600 python bench.py --seqs $seqs --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-widen --no-latency > gpurun_out/ab_align_${lib}_$seqs.json 2> gpurun_out/ab_align_${lib}_$seqs.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/ab_align_${lib}_$seqs.json").read().strip().splitlines()[-1])
print("$lib seqs $seqs value %.0f ms/step %.4f align %.4f search %.4f geom %.4f" % (d["value"], d["ms_per_step"], d["stages"]["sparse_align"]["ms"], d["stages"]["seeds_search"]["ms"], d["stages"]["seeds_geom"]["ms"]))
PY
  done
done
