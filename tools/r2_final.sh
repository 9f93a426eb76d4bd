#!/bin/bash
# end-of-round evidence of ONE build on one GPU: tests, smoke, bench arms and regimes (tools/r2_full.sh), then the ncu launch list and
# full capture of the same command (profiles/capture.sh), then the FAST capture.  TAG=$1
TAG=${1:-r2c}
bash tools/r2_full.sh $TAG
timeout 1200 bash profiles/capture.sh $TAG 4096 2>&1 | tail -2
python tools/fast_probe.py 1024 > gpurun_out/${TAG}_fast_probe.txt 2>&1; cat gpurun_out/${TAG}_fast_probe.txt
ncu --set full --clock-control none --import-source on -k regex:fast_kernel --launch-skip 2 -c 1 -f -o gpurun_out/${TAG}_fast python tools/fast_probe.py 256 > gpurun_out/${TAG}_ncu_fast.log 2>&1
ncu -i gpurun_out/${TAG}_fast.ncu-rep --page raw --csv > gpurun_out/${TAG}_fast_raw.csv
rm -f gpurun_out/${TAG}_fast.ncu-rep
python tools/pyr_probe.py 4096 2>&1 | tail -1
ls -la gpurun_out | grep ${TAG} | wc -l
