"""FAST detector probe for ncu: detect on a batch of rendered keyframes (python tools/fast_probe.py [seqs])."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from android_svo_b200 import capi, synth, frontend
seqs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = capi.Context(0)
cfg = synth.CONFIGS["C2"]
cam = capi.Camera.make(cfg["w"], cfg["h"], cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"])
tex = synth.make_texture(bench.TEX_SIZE)
d_tex = ctx.dev_alloc(tex.nbytes); ctx.dev_upload(d_tex, tex)
poses = bench.poses_for(range(seqs), (0,))
d_img = ctx.dev_alloc(seqs * cfg["w"] * cfg["h"])
ctx.synth_render(d_tex, bench.TEX_SIZE, bench.PPM, bench.PLANE_Z, cam, poses[:, 0], d_img)
ctx.frame_create(5, seqs, cfg["w"], cfg["h"], cfg["n_levels"])
ctx.frame_bind(5, d_img, cfg["w"])
for cell, thr in ((40, 20.0), (20, 10.0)):
    for rep in range(3):
        ctx.sync(); ctx.timer_start()
        cells, counts = ctx.fast_detect(5, cfg["n_pyr"], cell, thr)
        ms = ctx.timer_stop_ms()
    print("cell %d: %.3f ms per %d frames = %.2f us/frame, %.1f features/frame" % (cell, ms, seqs, ms / seqs * 1e3, counts.mean()))
