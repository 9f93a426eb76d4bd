"""Run as a script by tests/test_gpu_parity.py::test_pyramid_tma_kernel_bit_exact with SVOB200_PYRAMID_TMA set in the environment
(the library reads the knob once per process): the TMA-load pyramid kernel (cp.async.bulk.tensor + mbarrier, persistent CTAs) on
batches of 64+ device-resident frames against the oracle's vk::halfSample (vision.cpp:20-110), every level, bit for bit."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from android_svo_b200 import capi
from oracle import pyoracle
import scenes

assert os.environ.get("SVOB200_PYRAMID_TMA"), "set SVOB200_PYRAMID_TMA"
ctx = capi.Context(0)
oracle = pyoracle.Oracle()
checked = 0
for (h, w, n_levels, b, modes) in ((480, 640, 4, 96, None), (480, 752, 5, 80, None), (1080, 1920, 5, 64, None), (480, 640, 4, 70, [0, 0, 0]),
                                   (480, 640, 3, 65, [1, 0]), (66, 144, 4, 64, None), (64, 64, 7, 64, None), (480, 640, 2, 64, None)):
    base = [scenes.noise_image(h, w, s, blur=(s % 2 == 0)) for s in range(4)]
    imgs = np.stack([base[i % 4] if i % 5 else np.roll(base[i % 4], i, axis=1) for i in range(b)])
    d = ctx.dev_alloc(imgs.nbytes)
    ctx.dev_upload(d, imgs)
    ctx.frame_create(7, b, w, h, n_levels)
    ctx.frame_bind(7, d, w, round_modes=modes)
    ctx.sync()
    for im in (0, 1, 5, b // 2, b - 1):
        po = oracle.pyramid(imgs[im], n_levels, modes)
        for l in range(1, n_levels):
            got = ctx.frame_download(7, im, l)
            assert np.array_equal(got, po[l]), "%dx%d batch %d image %d level %d: %d px differ" % (w, h, b, im, l, (got != po[l]).sum())
            checked += 1
    ctx.frame_release(7)
    ctx.dev_free(d)
print("TMA pyramid OK: %d level images" % checked)
