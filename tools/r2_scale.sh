#!/bin/bash
# multi-GPU line (torchrun, one rank per GPU): bash tools/r2_scale.sh N TAG
N=$1; TAG=${2:-r2a}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 \
   --no-latency --no-widen --no-cpu-baseline > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
echo "rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], {k: v["ms"] for k, v in d["stages"].items() if v["ms"] > 0.02})
PY
