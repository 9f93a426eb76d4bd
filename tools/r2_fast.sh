#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -n "passed\|failed\|keyframe insertion" gpurun_out/pytest_gpu.log | tail -5; tail -5 gpurun_out/pytest_gpu.log
python tools/fast_probe.py 1024 > gpurun_out/fast_probe.txt 2>&1; cat gpurun_out/fast_probe.txt
ncu --set full --clock-control none --import-source on -k regex:fast_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r2b_fast python tools/fast_probe.py 256 > gpurun_out/ncu_fast.log 2>&1
ncu -i gpurun_out/r2b_fast.ncu-rep --page raw --csv > gpurun_out/r2b_fast_raw.csv
ncu -i gpurun_out/r2b_fast.ncu-rep --page source --csv > gpurun_out/r2b_fast_source.csv 2>/dev/null
rm -f gpurun_out/r2b_fast.ncu-rep
