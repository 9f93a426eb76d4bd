"""Per-phase cycle counts of the sparse-align kernel (needs the -DALIGN_TIMING build: SVOB200_LIB=.../libsvob200_T.so)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from android_svo_b200 import capi, synth
for name, seqs in (("C2", 1), ("C2", 512), ("C2", 4096)):
    ctx = capi.Context(0)
    cfg = synth.CONFIGS[name]
    wl = bench.GpuWorkload(ctx, capi, cfg, list(range(seqs)), cfg_name=name)
    order = bench.ping_pong(len(bench.POOL_INDICES), 20)
    for k in range(4):
        wl.step(order, k, capi.MEM_DEVICE)
    ctx.sync()
    # read the align results of the last step from the tracker: re-run the align through the API on the same frames is complex;
    # the tracker keeps d_align: expose through stats? use the API call instead on a fresh problem
    ar = np.zeros(seqs, capi.align_result_dt)
    ctx._ck(ctx.L.svob200_tracker_debug_align(wl.trk.h, capi._ptr(ar)))
    H = ar["H"][:, :6]
    it = ar["iters"].sum(1)
    sel = ar["H"][:, 6] / it > 0.9 * np.median(ar["H"][:, 6] / it)       # the probe's feature was visible throughout
    print(name, seqs, "iters mean %.1f" % it.mean(), "cycles: precompute %.0f  pass+reduce %.0f  solve %.0f  chain %.0f  decide+update %.0f  total %.0f" % tuple(H.mean(0)),
          "| per iteration: pass %.0f solve %.0f chain %.0f update %.0f" % tuple(H.mean(0)[1:5] / it.mean()),
          "| pass sub-phases per iteration (probe thread; problems where it worked in every iteration): project %.0f window %.0f pixels %.0f jres %.0f vote %.0f reduce %.0f sync1 %.0f total+sync2 %.0f (n=%d)"
          % (tuple((ar["H"][:, 6:14] / it[:, None])[sel].mean(0)) + (int(sel.sum()),)))
    wl.close(); ctx.close()
