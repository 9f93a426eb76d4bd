"""The C++ drop-in (android_svo_b200/host/svo_b200_dropin.cpp) against the reference itself.

oracle/ref_harness.cpp is a C++ program written against the reference's own headers and operator
surface (vk::halfSample, FastDetector::detect, SparseImgAlign::run, feature_alignment::align2D/1D,
warp::*, Matcher::findMatchDirect / findEpipolarMatchDirect, DepthFilter::addFrame -> updateSeeds).
It is linked twice: over the reference's unmodified sources (oracle/_ref/libsvo_ref.so) and over the
drop-in, where the hot-path symbols are OURS and run in CUDA (oracle/_ref/libsvo_dropin.so).  These
tests call the same harness entry points on both and compare: integer/byte outputs bit-exact, floats
within the north-star tolerances (pose 1e-4, pixel 1e-3 px, seeds 1e-5 relative).

The CPU test only checks that the drop-in links and resolves the reference's mangled symbols.
"""
import os
import subprocess
import numpy as np
import pytest

from android_svo_b200 import synth, frontend
from oracle.pyoracle import Ref, Cam, RefSeq, OUT
import scenes
from test_pipeline import make_sequence

DROPIN = os.path.join(OUT, "libsvo_dropin.so")

# the reference's hot-path symbols the drop-in must define (mangled names as the reference's callers bind them)
HOT_SYMBOLS = [
    "_ZN2vk10halfSampleERKN2cv3MatERS1_",
    "_ZN2vk14shiTomasiScoreERKN2cv3MatEii",
    "_ZN3svo17feature_alignment7align2DERKN2cv3MatEPhS5_iRN5Eigen6MatrixIdLi2ELi1ELi0ELi2ELi1EEEb",
    "_ZN3svo17feature_alignment7align1DERKN2cv3MatERKN5Eigen6MatrixIfLi2ELi1ELi0ELi2ELi1EEEPhSA_iRNS6_IdLi2ELi1ELi0ELi2ELi1EEERd",
    "_ZN3svo14SparseImgAlign3runESt10shared_ptrINS_5FrameEES3_",
    "_ZN3svo14SparseImgAlign20getFisherInformationEv",
    "_ZN3svo7Matcher15findMatchDirectERKNS_5PointERKNS_5FrameERN5Eigen6MatrixIdLi2ELi1ELi0ELi2ELi1EEE",
    "_ZN3svo7Matcher23findEpipolarMatchDirectERKNS_5FrameES3_RKNS_7FeatureEdddRd",
    "_ZN3svo17feature_detection12FastDetector6detectEPNS_5FrameERKSt6vectorIN2cv3MatESaIS6_EEdRNSt7__cxx114listIPNS_7FeatureESaISE_EEE",
    "_ZN3svo11Reprojector12reprojectMapESt10shared_ptrINS_5FrameEERSt6vectorISt4pairIS3_mESaIS6_EE",
    "_ZN3svo14pose_optimizer19optimizeGaussNewtonEdmbRSt10shared_ptrINS_5FrameEERdS5_S5_Rm",
    "_ZN3svo15B200DepthFilter11updateSeedsESt10shared_ptrINS_5FrameEE",
]


def test_dropin_defines_the_reference_symbols():
    if not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/libsvo_dropin.so not built (needs /root/reference)")
    out = subprocess.check_output(["nm", "-D", "--defined-only", DROPIN], text=True)
    defined = {line.split()[-1] for line in out.splitlines() if line.strip()}
    missing = [s for s in HOT_SYMBOLS if s not in defined]
    assert not missing, "drop-in does not define: %s" % missing
    und = subprocess.check_output(["nm", "-D", "--undefined-only", DROPIN], text=True)
    assert "svob200_sparse_align" in und and "svob200_seeds_update" in und, "drop-in does not go through the C ABI"
    ref_lib = os.path.join(OUT, "libsvo_ref.so")
    if os.path.exists(ref_lib):   # same mangled names in the reference build (except our subclass)
        rdef = {line.split()[-1] for line in subprocess.check_output(["nm", "-D", "--defined-only", ref_lib], text=True).splitlines() if line.strip()}
        assert all(s in rdef for s in HOT_SYMBOLS[:-1])


@pytest.fixture(scope="module")
def both(ref):
    d = Ref(dropin=True)
    if not d.available():
        pytest.fail("oracle/_ref/libsvo_dropin.so missing on a GPU box: build() must produce it where /root/reference exists")
    assert d.is_dropin() and not ref.is_dropin()
    return ref, d


@pytest.mark.gpu
def test_half_sample_and_pyramid(both):
    ref, d = both
    for (h, w) in ((480, 640), (480, 752), (270, 480), (60, 94)):
        img = scenes.noise_image(h, w, 7 + w)
        assert np.array_equal(ref.half_sample(img), d.half_sample(img))
    img = scenes.noise_image(480, 752, 3)
    pr, pd = ref.pyramid(img, 5), d.pyramid(img, 5)
    for l in range(5):
        assert np.array_equal(pr[l], pd[l]), "level %d" % l
    assert d.dropin_launches() > 0


@pytest.mark.gpu
def test_shi_tomasi_and_fast_detect(both):
    ref, d = both
    cfg, poses, imgs = scenes.scene("C2", 3)
    cam = scenes.cam_of(cfg, Cam)
    img = imgs[0]
    for (u, v) in ((100, 100), (5, 5), (320, 240), (634, 474), (635, 100), (17, 333)):
        assert ref.shi_tomasi(img, u, v) == d.shi_tomasi(img, u, v)
    for r in (ref, d):
        r.config(cfg["n_pyr"], cfg["max_level"], cfg["min_level"])
    for cell, thr, occ in ((40, 20.0, None), (20, 10.0, np.array([[100.5, 100.5], [333.0, 222.0]]))):
        pr, lr = ref.fast_detect(img, cam, cfg["n_pyr"], cell, thr, occ)
        pd, ld = d.fast_detect(img, cam, cfg["n_pyr"], cell, thr, occ)
        assert len(pr) > 50
        assert np.array_equal(pr, pd) and np.array_equal(lr, ld)


@pytest.mark.gpu
def test_sparse_align(both, oracle):
    ref, d = both
    cfg, poses, imgs, kf, last_px = make_sequence(oracle, "C2", 0x00C0FFEE, 4)
    cam = scenes.cam_of(cfg, Cam)
    for r in (ref, d):
        r.config(cfg["n_pyr"], cfg["max_level"], cfg["min_level"])
    pts = kf["pt_world"].copy()
    pts[5, 0] = np.nan                      # a feature without a point (sparse_img_align.cpp:118)
    for k in (1, 2):
        a = ref.sparse_align(imgs[k - 1], imgs[k], cam, cfg["max_level"], cfg["min_level"], 30, poses[k - 1], poses[k - 1], last_px[k - 1], pts)
        b = d.sparse_align(imgs[k - 1], imgs[k], cam, cfg["max_level"], cfg["min_level"], 30, poses[k - 1], poses[k - 1], last_px[k - 1], pts)
        assert a["n_tracked"] == b["n_tracked"] and a["n_meas"] == b["n_meas"] and a["stop"] == b["stop"]
        assert np.array_equal(a["iters"], b["iters"]), "Gauss-Newton evaluations per level differ"
        rot, trans = synth.pose_error(a["T_cur_w"], b["T_cur_w"])
        assert rot <= 1e-4 and trans <= 2e-4        # north-star tolerance (1e-4 rad, 1e-4 x scene scale 2 m)
        assert rot < 1e-9 and trans < 1e-9          # what is actually achieved
        assert np.allclose(a["H"], b["H"], rtol=1e-9) and np.allclose(a["Jres"], b["Jres"], rtol=1e-7, atol=1e-7)
        assert abs(a["chi2"] - b["chi2"]) <= 1e-5 * abs(a["chi2"])


@pytest.mark.gpu
def test_warp_align_and_match_direct(both, oracle):
    ref, d = both
    cfg, poses, imgs, kf, last_px = make_sequence(oracle, "C2", 0x00C0FFEE + 1, 4)
    cam = scenes.cam_of(cfg, Cam)
    for r in (ref, d):
        r.config(cfg["n_pyr"], cfg["max_level"], cfg["min_level"])
    rng = np.random.RandomState(5)
    n_ok = 0
    for i in range(0, 40):
        px0, lvl, pt = kf["kf_px"][i], int(kf["kf_level"][i]), kf["pt_world"][i]
        guess = frontend.project_many(cfg, poses[2], pt[None])[0] + rng.uniform(-1.5, 1.5, 2)
        kw = dict(imgs=[imgs[0], imgs[2]], T_f_w=np.stack([poses[0], poses[2]]), cam=cam, pt_world=pt, obs_frame=[0], obs_px=px0,
                  obs_level=[lvl], cur_frame=1, px_cur=guess)
        a, b = ref.find_match_direct(**kw), d.find_match_direct(**kw)
        assert a["success"] == b["success"] and a["chosen"] == b["chosen"] and a["search_level"] == b["search_level"]
        assert np.array_equal(a["pwb"], b["pwb"]) and np.array_equal(a["patch"], b["patch"])      # warped patches bit-exact
        assert np.array_equal(a["A"], b["A"])
        assert np.abs(a["px_cur"] - b["px_cur"]).max() == 0.0                                   # LK replays the float chains: exact
        n_ok += a["success"]
        # the stand-alone pieces: warp::* and align2D / align1D on the same patch
        f_ref = ref.cam2world(cam, px0[0], px0[1])
        depth = np.linalg.norm(pt - synth.se3_inverse(poses[0])[:3])
        T_cur_ref = ref.se3_mul(poses[2], ref.se3_inverse(poses[0]))
        pyr0 = oracle.pyramid(imgs[0], cfg["n_levels"])
        wa = ref.warp(pyr0[lvl], cam, px0, f_ref, depth, T_cur_ref, lvl, cfg["n_pyr"] - 1)
        wb = d.warp(pyr0[lvl], cam, px0, f_ref, depth, T_cur_ref, lvl, cfg["n_pyr"] - 1)
        assert np.array_equal(wa[0], wb[0]) and wa[1] == wb[1] and np.array_equal(wa[2], wb[2])
        pyr2 = oracle.pyramid(imgs[2], cfg["n_levels"])
        L = a["search_level"]
        pa = ref.align2d(pyr2[L], a["pwb"], a["patch"], 10, guess / (1 << L))
        pb = d.align2d(pyr2[L], a["pwb"], a["patch"], 10, guess / (1 << L))
        assert pa[0] == pb[0] and np.array_equal(pa[1], pb[1])
        dirv = np.array([0.6, 0.8], np.float32)
        qa = ref.align1d(pyr2[L], dirv, a["pwb"], a["patch"], 10, guess / (1 << L))
        qb = d.align1d(pyr2[L], dirv, a["pwb"], a["patch"], 10, guess / (1 << L))
        assert qa[0] == qb[0] and np.array_equal(qa[1], qb[1]) and qa[2] == qb[2]
    assert n_ok > 30


@pytest.mark.gpu
@pytest.mark.parametrize("align_1d", [0, 1])
def test_epipolar_match(both, oracle, align_1d):
    ref, d = both
    cfg, poses, imgs, kf, last_px = make_sequence(oracle, "C2", 0x00C0FFEE + 2, 6)
    cam = scenes.cam_of(cfg, Cam)
    for r in (ref, d):
        r.config(cfg["n_pyr"], cfg["max_level"], cfg["min_level"])
    n_ok = 0
    for i in range(0, 60):
        px0, lvl = kf["seed_px"][i], int(kf["seed_level"][i])
        ranges = ((2.4, 1.2, 6.0), (2.0, 1.9, 2.1), (2.0, 1.999, 2.001), (2.2, 0.05, 1e8))
        d_est, d_min, d_max = ranges[i % 4]
        kw = dict(ref_img=imgs[0], cur_img=imgs[5], cam=cam, T_ref_w=poses[0], T_cur_w=poses[5], px_ref=px0, level_ref=lvl,
                  d_est=d_est, d_min=d_min, d_max=d_max, align_1d=align_1d)
        a, b = ref.find_epipolar_match(**kw), d.find_epipolar_match(**kw)
        assert a["success"] == b["success"] and a["search_level"] == b["search_level"] and a["reject"] == b["reject"]
        assert np.array_equal(a["pwb"], b["pwb"]) and np.array_equal(a["patch"], b["patch"])
        assert a["epi_length"] == b["epi_length"] and np.array_equal(a["A"], b["A"])
        assert np.abs(a["px_cur"] - b["px_cur"]).max() <= 1e-3
        if a["success"]:
            assert abs(a["depth"] - b["depth"]) <= 1e-9 * abs(a["depth"])
            n_ok += 1
    assert n_ok > 20


@pytest.mark.gpu
def test_update_seeds_list_semantics(both, oracle):
    """DepthFilter::addFrame -> (B200DepthFilter::)updateSeeds over a std::list<Seed>: erase / callback / b++ semantics."""
    ref, d = both
    cfg, poses, imgs, kf, last_px = make_sequence(oracle, "C2", 0x00C0FFEE + 3, 6)
    cam = scenes.cam_of(cfg, Cam)
    for r in (ref, d):
        r.config(cfg["n_pyr"], cfg["max_level"], cfg["min_level"])
    S = 300
    state = np.tile(ref.seed_init(2.4, 1.2), (S, 1))
    state[::7, 4] *= 1e-4          # nearly converged seeds -> converge on this frame (callback + erase)
    state[3, 2] = np.nan           # NaN mean -> erased through the z_inv_min branch
    sa = sb = state
    for k in (2, 4, 5):
        sa, sta = ref.update_seeds([imgs[0]], poses[0][None], imgs[k], poses[k], cam, np.zeros(S, np.int32), kf["seed_px"][:S], kf["seed_level"][:S], sa)
        sb, stb = d.update_seeds([imgs[0]], poses[0][None], imgs[k], poses[k], cam, np.zeros(S, np.int32), kf["seed_px"][:S], kf["seed_level"][:S], sb)
        assert np.array_equal(sta, stb), "seed list membership (kept / converged / erased) differs"
        assert (sta == 1).sum() > 10
        fin = np.isfinite(sa).all(axis=1)
        assert np.array_equal(fin, np.isfinite(sb).all(axis=1))
        # same poses and same input state on both sides: the drop-in must reproduce the reference's seed bits
        n_diff = int((sa[fin].view(np.uint32) != sb[fin].view(np.uint32)).any(axis=1).sum())
        assert n_diff == 0, "%d of %d seeds differ in bits under identical poses" % (n_diff, int(fin.sum()))
        sb = sa                      # continue both from the same state


@pytest.mark.gpu
def test_frontend_sequence(both, oracle):
    """The per-frame chain the reference runs (new Frame -> SparseImgAlign::run -> findMatchDirect per point ->
    DepthFilter::addFrame), every operator through the drop-in, frame after frame."""
    ref, d = both
    cfg, poses, imgs, kf, last_px = make_sequence(oracle, "C2", 0x00C0FFEE + 4, 11)
    cam = scenes.cam_of(cfg, Cam)
    args = (cam, cfg["n_levels"], cfg["max_level"], cfg["min_level"], cfg["n_pyr"])
    for r in (ref, d):
        r.config(cfg["n_pyr"], cfg["max_level"], cfg["min_level"])
    sr, sd = RefSeq(ref, *args), RefSeq(d, *args)
    try:
        for s in (sr, sd):
            s.set_keyframe(imgs[0], poses[0], kf["kf_px"], kf["kf_level"], kf["pt_world"], kf["seed_px"], kf["seed_level"])
            s.set_last(imgs[0])
        conv = n_outside = 0
        for k in range(1, 11):
            a, pxa, oka = sr.step(imgs[k], poses[k - 1], last_px[k - 1], want_px=True)
            b, pxb, okb = sd.step(imgs[k], poses[k - 1], last_px[k - 1], want_px=True)
            assert a.n_tracked == b.n_tracked and a.align_iters == b.align_iters
            rot, trans = synth.pose_error(np.array(a.T_cur_w[:]), np.array(b.T_cur_w[:]))
            assert rot <= 1e-4 and trans <= 2e-4 and rot < 1e-9 and trans < 1e-9
            assert a.n_matched == b.n_matched and np.array_equal(oka, okb)
            assert np.abs(pxa - pxb).max() <= 1e-3
            assert a.n_seeds_converged == b.n_seeds_converged
            xa, xb = sr.seeds(), sd.seeds()
            assert np.array_equal(xa[:, 0] < 0, xb[:, 0] < 0)
            # the drop-in's pose differs by ~1e-10 (alignment sums are tolerance-matched); seeds outside 1e-5 are counted —
            # tests/test_pipeline.py::SeedParity proves each such deviation is that pose difference and nothing else
            outside = ~np.isclose(xa, xb, rtol=1e-5, atol=0).all(axis=1)
            n_outside += int(outside.sum())
            assert np.allclose(xa, xb, rtol=2e-2, atol=0)
            conv += a.n_seeds_converged
        assert conv > 50
        print("drop-in sequence: %d of %d seed-frames outside 1e-5" % (n_outside, 10 * len(xa)))
        assert n_outside <= 0.01 * 10 * len(xa)
    finally:
        sr.close(); sd.close()


@pytest.mark.gpu
def test_frontend_sequence_two_threads(both, oracle):
    """The reference's own two-thread layout over the drop-in (DepthFilter::startThread, depth_filter.cpp:63-103): the tracking
    thread's operators and the depth-filter thread's updateSeeds drive TWO svob200 contexts (two streams) concurrently, nothing
    is paced except the frame queue.  Tracking results are compared frame by frame, the seeds after the queue has drained,
    against the reference TUs run sequentially (an update depends on its frame and the seed list only, so the order of the
    two threads does not change it).  No re-seeding: converged seeds leave the list, as in the app."""
    ref, d = both
    cfg, poses, imgs, kf, last_px = make_sequence(oracle, "C2", 0x00C0FFEE + 9, 11)
    cam = scenes.cam_of(cfg, Cam)
    args = (cam, cfg["n_levels"], cfg["max_level"], cfg["min_level"], cfg["n_pyr"])
    for r in (ref, d):
        r.config(cfg["n_pyr"], cfg["max_level"], cfg["min_level"])
    sr, sd = RefSeq(ref, *args, reseed=0), RefSeq(d, *args, reseed=0)
    try:
        sd.set_threaded()
        for s in (sr, sd):
            s.set_keyframe(imgs[0], poses[0], kf["kf_px"], kf["kf_level"], kf["pt_world"], kf["seed_px"], kf["seed_level"])
            s.set_last(imgs[0])
        for k in range(1, 11):
            a, pxa, oka = sr.step(imgs[k], poses[k - 1], last_px[k - 1], want_px=True)
            b, pxb, okb = sd.step(imgs[k], poses[k - 1], last_px[k - 1], want_px=True)
            assert a.n_tracked == b.n_tracked and a.align_iters == b.align_iters
            rot, trans = synth.pose_error(np.array(a.T_cur_w[:]), np.array(b.T_cur_w[:]))
            assert rot < 1e-9 and trans < 1e-9
            assert a.n_matched == b.n_matched and np.array_equal(oka, okb)
            assert np.abs(pxa - pxb).max() <= 1e-3
        sd.drain()
        xa, xb = sr.seeds(), sd.seeds()
        assert np.array_equal(xa[:, 0] < 0, xb[:, 0] < 0), "different seeds finished"
        assert (xa[:, 0] < 0).sum() > 50, "the sequence should converge a good part of its seeds"
        outside = ~np.isclose(xa, xb, rtol=1e-5, atol=0).all(axis=1)
        assert np.allclose(xa, xb, rtol=2e-2, atol=0)
        print("two-thread drop-in sequence: %d of %d seeds outside 1e-5 after 10 frames" % (int(outside.sum()), len(xa)))
        assert outside.sum() <= 0.01 * len(xa)
    finally:
        sr.close(); sd.close()


# ---------------------------------------------------------------- callers either side of the hot path
@pytest.fixture(scope="module")
def both_map(both):
    from oracle import pyoracle_map as pm
    ref, d = both
    return pm.RefMap(ref), pm.RefMap(d)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,max_fts", [(5, 120), (7, 12)])
def test_reprojector_dropin(both_map, oracle, seed, max_fts):
    """svo::Reprojector::reprojectMap of the drop-in (device grid + batched matching) against the reference's, on the same
    Map built by the same harness: counters, per-point side effects, the new features of the frame (bit-exact pixels)"""
    import map_scenes as ms
    rm, dm = both_map
    sc = ms.build_map_scene(oracle, seed=seed)
    cfg = sc["cfg"]
    cam = Cam.make(cfg["w"], cfg["h"], cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"])
    out = []
    for m in (rm, dm):
        m.config(cfg["n_pyr"], sc["cell"], max_fts)
        out.append(m.reproject_map(sc["kf_imgs"], sc["T_kf"], sc["cur_img"], sc["T_cur"], cam, sc["points"], sc["obs"], sc["n_candidates"]))
    a, b = out
    assert (a["n_matches"], a["n_trials"]) == (b["n_matches"], b["n_trials"]) and a["n_matches"] > 10
    for k in ("n_failed", "n_succeeded", "type_after", "new_point", "new_level", "new_type", "overlap_kf", "overlap_cnt"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["new_px"], b["new_px"])
    assert np.allclose(a["new_grad"], b["new_grad"], rtol=0, atol=1e-12)


@pytest.mark.gpu
def test_pose_optimizer_dropin(both_map):
    import map_scenes as ms
    rm, dm = both_map
    for seed in (3, 9):
        s = ms.pose_opt_scene(seed=seed)
        cfg = s["cfg"]
        cam = Cam.make(cfg["w"], cfg["h"], cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"])
        img = np.zeros((cfg["h"], cfg["w"]), np.uint8)
        a = rm.pose_optimize(cam, img, s["px"], s["level"], s["pos"], s["T_init"])
        b = dm.pose_optimize(cam, img, s["px"], s["level"], s["pos"], s["T_init"])
        rot, trans = synth.pose_error(a["T"], b["T"])
        assert rot <= 1e-9 and trans <= 1e-9                    # spec: 1e-4 rad / 1e-4 of scene scale
        assert np.array_equal(a["outlier"], b["outlier"]) and a["num_obs"] == b["num_obs"]
        assert a["estimated_scale"] == b["estimated_scale"] and a["error_init"] == b["error_init"]
        assert np.isclose(a["error_final"], b["error_final"], rtol=1e-9)
        assert np.allclose(a["A"], b["A"], rtol=1e-6, atol=1e-6 * np.abs(a["A"]).max())


@pytest.mark.gpu
def test_optimize_structure_dropin(both_map):
    """FrameHandlerBase::optimizeStructure: the selection by last_structure_optim_ and every Point::optimize, bit-exact"""
    import map_scenes as ms
    rm, dm = both_map
    cfg = ms.SMALL
    cam = Cam.make(cfg["w"], cfg["h"], cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"])
    img = np.zeros((cfg["h"], cfg["w"]), np.uint8)
    pts = ms.point_opt_scene()
    off = np.concatenate([[0], np.cumsum([len(p["T"]) for p in pts])])
    T = np.concatenate([p["T"] for p in pts]); f = np.concatenate([p["f"] for p in pts])
    pos0 = np.stack([p["pos0"] for p in pts])
    last = np.arange(len(pts))[::-1]
    a_pos, a_done = rm.optimize_structure(cam, img, off, T, f, pos0, last, 20, 5)
    b_pos, b_done = dm.optimize_structure(cam, img, off, T, f, pos0, last, 20, 5)
    assert a_done.sum() == 20 and np.array_equal(a_done, b_done)
    assert np.array_equal(a_pos, b_pos)
