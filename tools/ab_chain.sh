# GPU box: chain property + alignment parity tests, the per-phase cycle split (timing build) and the latency legs of the bench
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_pipeline.py -m gpu -x -q -k "chi2_chain or sparse_align or tracker or pipeline" 2>&1 | tail -6
SVOB200_LIB=$PWD/android_svo_b200/lib/libsvob200_T.so timeout 250 python tools/align_timing.py 2>&1 | tail -4
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-widen > gpurun_out/chain_C.json 2> gpurun_out/chain_C.err
python -c "
import json; d=json.load(open('gpurun_out/chain_C.json')); print(d['value'], d['ms_per_step']); print({k:(v['resident']['p50_ms'],v['resident']['p95_ms'],v['stages_ms_resident']['sparse_align']) for k,v in d['latency'].items()}); print({k:v['ms'] for k,v in d['stages'].items()})"
