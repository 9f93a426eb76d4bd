#!/bin/bash
# ncu evidence for one round, run on the GPU box from the repo root (under gpurun):
#   bash profiles/capture.sh r1 [SEQS]
# 1. launch list  (gpu__time_duration.sum, --clock-control none)  -> gpurun_out/<tag>_launches.csv
# 2. full capture of the hot kernels of two tracker steps      -> gpurun_out/<tag>_full.ncu-rep (+ raw csv)
# Both only after the same command has exited 0 without ncu.  Numbers printed under ncu are never bench values.
set -e
TAG=${1:-r1}
SEQS=${2:-4096}
CMD="python bench.py --seqs $SEQS --steps 2 --warmup 1 --no-latency --no-cpu-baseline --no-e2e --no-widen"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'pyramid_fused|sparse_align|match_prepare|seeds_geom|epi_search|lk_refine|seeds_finish' \
    --launch-skip 8 -c 16 -f -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv
echo "capture done: $(wc -l < gpurun_out/${TAG}_launches.csv) launch rows"
