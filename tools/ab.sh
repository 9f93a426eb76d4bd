#!/bin/bash
# A/B runs of alternative builds of libsvob200.so (android_svo_b200/lib/libsvob200_<tag>.so) on the GPU box:
#   bash tools/ab.sh A B5 B6
for v in "$@"; do
  SVOB200_LIB=$PWD/android_svo_b200/lib/libsvob200_$v.so python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/ab_$v.json')); print('$v', d['value'], d['ms_per_step']); print({k:v['ms'] for k,v in d['stages'].items()})"
  tail -2 gpurun_out/ab_$v.err
done
