# GPU box: A/B of library builds over the resident bench (android_svo_b200/lib/libsvob200_<v>.so; "main" = libsvob200.so)
mkdir -p gpurun_out
for v in ${AB_VARIANTS:-B main B main}; do
  L=$PWD/android_svo_b200/lib/libsvob200_$v.so; [ $v = main ] && L=$PWD/android_svo_b200/lib/libsvob200.so
  SVOB200_LIB=$L timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-widen ${AB_FLAGS:---no-latency} > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/ab_$v.json')); print('$v', d['value'], d['ms_per_step'], 'align', d['stages']['sparse_align']['ms'], 'clk', d['clocks']['sm_mhz'])"
done
