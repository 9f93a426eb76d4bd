set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for v in A B; do
  L=$PWD/android_svo_b200/lib/libsvob200.so; [ $v = A ] && L=$PWD/android_svo_b200/lib/libsvob200_A.so
  SVOB200_LIB=$L timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-widen > gpurun_out/chain_$v.json 2> gpurun_out/chain_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/chain_$v.json')); print('$v', d['value'], d['ms_per_step']); print(d.get('latency')); print({k:v['ms'] for k,v in d['stages'].items()})"
done
