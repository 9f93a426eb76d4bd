#!/bin/bash
mkdir -p gpurun_out
Q="--steps 30 --warmup 5 --no-latency --no-cpu-baseline --no-widen --no-e2e"
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py $Q $EXTRA > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_$name.json").read().strip().splitlines()[-1])
    print("$name: value", d["value"], "ms/step", d["ms_per_step"], {k: v["ms"] for k, v in d["stages"].items() if k in ("seeds_geom", "seeds_finish", "seeds_search", "frame+pyramid")})
except Exception as e:
    print("$name failed", e, open("gpurun_out/ab_$name.err").read()[-400:])
PY
}
for S in 512 1024 4096; do
  EXTRA="--seqs $S" run s${S}_w1 SVOB200_SEARCH_WAVES=1
  EXTRA="--seqs $S" run s${S}_w2 SVOB200_SEARCH_WAVES=2
  EXTRA="--seqs $S" run s${S}_w4 SVOB200_SEARCH_WAVES=4
done
EXTRA="--seqs 4096" run s4096_g8 SVOB200_LIB=$PWD/android_svo_b200/lib/libsvob200_g8.so
EXTRA="--seqs 512" run s512_g8 SVOB200_LIB=$PWD/android_svo_b200/lib/libsvob200_g8.so
