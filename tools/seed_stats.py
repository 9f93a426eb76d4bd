"""Workload accounting for the depth-filter search kernel: distribution of epipolar walk lengths / ZMSSD evaluations
per seed in the bench's steady state (run on the GPU box): python tools/seed_stats.py [seqs]"""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from android_svo_b200 import capi, synth

seqs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = capi.Context(0)
cfg = synth.CONFIGS[bench.CFG_NAME]
wl = bench.GpuWorkload(ctx, capi, cfg, list(range(seqs)))
order = bench.ping_pong(len(bench.POOL_INDICES), 40)
out = {}
for k in range(12):
    wl.step(order, k, capi.MEM_DEVICE)
    if k in (0, 3, 11):
        obs = wl.trk.seed_obs()
        ne, el, st = obs["n_evals"], obs["epi_length"], obs["status"]
        act = st >= 3
        h = np.bincount(np.minimum(ne[act], 64), minlength=65)
        out["step%d" % k] = dict(active_frac=float(act.mean()), status_hist=np.bincount(st, minlength=7).tolist(),
                                 n_evals_mean_all=float(ne.mean()), n_evals_mean_active=float(ne[act].mean()),
                                 n_evals_pct=[float(np.percentile(ne[act], p)) for p in (50, 75, 90, 95, 99, 100)],
                                 epi_len_pct=[float(np.percentile(el[act], p)) for p in (50, 75, 90, 95, 99, 100)],
                                 frac_direct=float((el[act] < 2.0).mean()), hist_0_8=h[:9].tolist(), hist_ge64=int(h[64]),
                                 sum_evals_share_of_long=float(ne[act][ne[act] > 24].sum() / max(ne[act].sum(), 1)))
print(json.dumps(out, indent=1))
