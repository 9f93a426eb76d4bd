#!/bin/bash
# the epipolar search kernel in the young-seed regime (mean 26 ZMSSD evaluations per seed): one full ncu capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dropin.py -q -m gpu 2>&1 | tail -2
CMD="python bench.py --seqs 4096 --steps 2 --warmup 3 --seed-regime young --no-latency --no-cpu-baseline --no-e2e --no-widen"
$CMD > gpurun_out/r2c_young_plain.json 2> gpurun_out/r2c_young_plain.err; echo "plain rc=$?"
ncu --set full --clock-control none --import-source on -k regex:epi_search --launch-skip 4 -c 1 -f -o gpurun_out/r2c_young_search $CMD > gpurun_out/r2c_young_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2c_young_search.ncu-rep --page raw --csv > gpurun_out/r2c_young_search_raw.csv
ncu -i gpurun_out/r2c_young_search.ncu-rep --page source --csv --print-source sass > gpurun_out/r2c_young_search_sass.csv 2>/dev/null
rm -f gpurun_out/r2c_young_search.ncu-rep
wc -l gpurun_out/r2c_young_search_raw.csv gpurun_out/r2c_young_search_sass.csv
