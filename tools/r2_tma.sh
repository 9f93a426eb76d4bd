#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/pyr_probe.py 4096 gpurun_out/pyr_base.npz 2>&1 | tail -2
for cfg in "6 0" "5 0" "7 0" "6 1" "5 1" "7 1" "4 1"; do
  set -- $cfg
  SVOB200_PYRAMID_TMA=$1 SVOB200_PYRAMID_TMA_WIDE=$2 timeout 120 python tools/pyr_probe.py 4096 gpurun_out/pyr_tma$1_$2.npz 2>&1 | tail -1 | sed "s/^/wide=$2 /"
done
python - <<'PY'
import numpy as np, glob
a = np.load("gpurun_out/pyr_base.npz")
for fn in sorted(glob.glob("gpurun_out/pyr_tma*_*.npz")):
    b = np.load(fn)
    bad = [k for k in a.files if not np.array_equal(a[k], b[k])]
    print(fn, "%d level images compared, %d differ %s" % (len(a.files), len(bad), bad[:4]))
PY
