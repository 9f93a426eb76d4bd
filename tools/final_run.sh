# GPU box, end of round: GPU tests, smoke, the two bench arms, then the ncu evidence of the same build (tag = $1)
TAG=${1:-r1e}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; tail -c 600 gpurun_out/${TAG}_ref.json
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')}); print(d['roofline']); print({k:(v['resident']['p50_ms'],v['host_buffers']['p50_ms']) for k,v in d['latency'].items()}); print(d['cpu_baseline'])"
timeout 1200 bash profiles/capture.sh $TAG 4096 2>&1 | tail -2
