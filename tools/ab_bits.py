"""A/B bit-identity of the sparse alignment between two builds of the library:
    SVOB200_LIB=.../libsvob200_old.so python tools/ab_bits.py gpurun_out/ab_old.npz
    python tools/ab_bits.py gpurun_out/ab_new.npz
    python tools/ab_bits.py --compare gpurun_out/ab_old.npz gpurun_out/ab_new.npz
Every alignment result record (pose, H, Jres, x, chi2, iteration counts, exact-chain count) of every step is kept."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if sys.argv[1] == "--compare":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    bad = 0
    for k in a.files:
        same = all(a[k][n].tobytes() == b[k][n].tobytes() for n in a[k].dtype.names)      # named fields only: the records have padding
        if not same:
            bad += 1
            ra, rb = a[k], b[k]
            diff = [n for n in ra.dtype.names if ra[n].tobytes() != rb[n].tobytes()]
            print("DIFF", k, diff, "iters", ra["iters"].sum(), rb["iters"].sum())
    print("%d record sets compared, %d differ" % (len(a.files), bad))
    sys.exit(1 if bad else 0)
import bench
from android_svo_b200 import capi, synth
out = {}
for name, seqs in (("C2", 1), ("C2", 3), ("C2", 24), ("C2", 48), ("C2", 300), ("C2", 1024), ("C3", 1), ("C3", 40), ("C4", 1), ("C4", 6)):
    ctx = capi.Context(0)
    cfg = synth.CONFIGS[name]
    wl = bench.GpuWorkload(ctx, capi, cfg, list(range(seqs)), cfg_name=name)
    order = bench.ping_pong(len(bench.POOL_INDICES), 20)
    for k in range(6):
        wl.step(order, k, capi.MEM_DEVICE)
        ctx.sync()
        ar = np.zeros(seqs, capi.align_result_dt)
        ctx._ck(ctx.L.svob200_tracker_debug_align(wl.trk.h, capi._ptr(ar)))
        out["%s_%d_step%d" % (name, seqs, k)] = ar
    print(name, seqs, "iters mean %.2f" % ar["iters"].sum(1).mean(), "n_exact mean %.2f" % ar["n_exact_chi2"].mean())
    wl.close(); ctx.close()
np.savez(sys.argv[1], **out)
