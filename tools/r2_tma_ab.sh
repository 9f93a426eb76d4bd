#!/bin/bash
mkdir -p gpurun_out
for seqs in 2048 1024 512; do
for R in 0 5 0 5; do
  SVOB200_PYRAMID_TMA=$R timeout 600 python bench.py --seqs $seqs --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-widen --no-latency 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tma=$R seqs $seqs', d['value'], d['ms_per_step'], d['stages']['frame+pyramid']['ms'])"
done
done
