#!/bin/bash
# old vs new alignment kernel: bit-identity of every result record, then timing of both
mkdir -p gpurun_out
SVOB200_LIB=/root/repo/android_svo_b200/lib/libsvob200_old.so timeout 600 python tools/ab_bits.py gpurun_out/ab_old.npz > gpurun_out/ab_old.log 2>&1; echo "old rc=$?"; tail -3 gpurun_out/ab_old.log
timeout 600 python tools/ab_bits.py gpurun_out/ab_new.npz > gpurun_out/ab_new.log 2>&1; echo "new rc=$?"; tail -3 gpurun_out/ab_new.log
python tools/ab_bits.py --compare gpurun_out/ab_old.npz gpurun_out/ab_new.npz 2>&1 | tail -8
for lib in old new; do
  L=/root/repo/android_svo_b200/lib/libsvob200.so; [ $lib = old ] && L=/root/repo/android_svo_b200/lib/libsvob200_old.so
  for seqs in 4096 512; do
    SVOB200_LIB=$L timeout 600 python bench.py --seqs $seqs --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-widen > gpurun_out/ab_align_${lib}_$seqs.json 2> gpurun_out/ab_align_${lib}_$seqs.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/ab_align_${lib}_$seqs.json").read().strip().splitlines()[-1])
print("$lib seqs $seqs value %.0f ms/step %.4f align %.4f" % (d["value"], d["ms_per_step"], d["stages"]["sparse_align"]["ms"]), "C2 lat", d.get("latency",{}).get("C2",{}).get("resident",{}).get("p50_ms"), d.get("latency",{}).get("C2",{}).get("stages_ms_resident",{}).get("sparse_align"))
PY
  done
done
