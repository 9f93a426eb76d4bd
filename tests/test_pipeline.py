"""Per-frame front-end step (pyramid -> sparse align -> reprojection refinement -> seed update):
the C restatement against the real reference (CPU), and svob200_tracker_step against the C
restatement (GPU)."""
import numpy as np
import pytest

from android_svo_b200 import synth, frontend
from oracle.pyoracle import Cam, OracleSeq, RefSeq
import scenes


def recycle(cells, thr, n):
    """cells above thr, recycled when the view holds fewer than n (full-size configs: same load as the bench)"""
    good = cells[cells["score"].astype(np.float64) > thr]
    out = np.resize(good, n) if len(good) < n else good[:n]
    # keyframe_setup takes the first n cells above thr: hand it exactly those, with scores that pass
    out = out.copy(); out["score"] = np.maximum(out["score"], np.float32(thr + 1))
    return out


def make_sequence(oracle, name, seed, n_frames, tex_size=1024, stride=3, allow_recycle=False):
    cfg = synth.CONFIGS[name]
    poses = synth.trajectory(n_frames * stride, seed=seed)[::stride]
    tex = scenes.texture(tex_size)
    imgs = [synth.render(tex, cfg, T) for T in poses]
    fc, ft, sc, st = frontend.DETECT[name]
    pyr = oracle.pyramid(imgs[0], cfg["n_levels"])
    _, fcells = oracle.fast_detect(pyr, cfg["n_pyr"], fc, ft)
    _, scells = oracle.fast_detect(pyr, cfg["n_pyr"], sc, st)
    if allow_recycle:
        fcells, scells = recycle(fcells, ft, cfg["n_features"]), recycle(scells, st, cfg["n_seeds"])
    kf = frontend.keyframe_setup(cfg, poses[0], fcells, scells, ft, st)
    last_px = [frontend.project_many(cfg, poses[k], kf["pt_world"]) for k in range(n_frames)]
    return cfg, poses, imgs, kf, last_px


def test_oracle_step_matches_reference(oracle, ref):
    cfg, poses, imgs, kf, last_px = make_sequence(oracle, "C2", 0x00C0FFEE, 11)
    cam = scenes.cam_of(cfg, Cam)
    args = (cam, cfg["n_levels"], cfg["max_level"], cfg["min_level"], cfg["n_pyr"])
    so, sr = OracleSeq(oracle, *args), RefSeq(ref, *args)
    try:
        for s in (so, sr):
            s.set_keyframe(imgs[0], poses[0], kf["kf_px"], kf["kf_level"], kf["pt_world"], kf["seed_px"], kf["seed_level"])
            s.set_last(imgs[0])
        total_conv = 0
        for k in range(1, 11):
            a, pxa, oka = so.step(imgs[k], poses[k - 1], last_px[k - 1], want_px=True)
            b, pxb, okb = sr.step(imgs[k], poses[k - 1], last_px[k - 1], want_px=True)
            assert a.n_tracked == b.n_tracked and a.align_iters == b.align_iters
            rot, trans = synth.pose_error(np.array(a.T_cur_w[:]), np.array(b.T_cur_w[:]))
            assert rot < 1e-12 and trans < 1e-12
            assert a.n_matched == b.n_matched and np.array_equal(oka, okb)
            assert np.abs(pxa - pxb).max() < 1e-9
            assert a.n_seeds_converged == b.n_seeds_converged
            sa, sb = so.seeds(), sr.seeds()
            # float seed state: one ulp of difference in 1/z (pose differs by ~1e-16 through the LDL^T) is amplified
            # by the cancellation in sigma2 = C1*(s2+m^2) + C2*(...) - mu_new^2, so: nearly all exact, all close
            assert np.isclose(sa, sb, rtol=1e-6, atol=0).all(axis=1).mean() > 0.99, "seed states differ after frame %d" % k
            assert np.allclose(sa, sb, rtol=1e-3, atol=0)
            total_conv += a.n_seeds_converged
            # sanity: alignment recovers the true pose to a few mm (stops at level 2)
            grot, gtrans = synth.pose_error(np.array(a.T_cur_w[:]), poses[k])
            assert grot < 1e-2 and gtrans < 3e-2
            assert a.n_tracked > 90 and a.n_matched > 90
        assert total_conv > 150
    finally:
        so.close(); sr.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,batch,n_frames", [("C2", 3, 10), ("C3", 2, 4)])
def test_tracker_step_matches_oracle(ctx, oracle, name, batch, n_frames):
    from android_svo_b200 import capi
    seqs = [make_sequence(oracle, name, 0x00C0FFEE + i, n_frames + 1) for i in range(batch)]
    cfg = seqs[0][0]
    cam_o, cam_g = scenes.cam_of(cfg, Cam), scenes.cam_of(cfg, capi.Camera)
    args = (cfg["n_levels"], cfg["max_level"], cfg["min_level"], cfg["n_pyr"])
    oseqs = [OracleSeq(oracle, cam_o, *args) for _ in range(batch)]
    trk = capi.Tracker(ctx, cam_g, batch, *args)
    try:
        N, S = cfg["n_features"], cfg["n_seeds"]
        for s, (_, poses, imgs, kf, _) in zip(oseqs, seqs):
            s.set_keyframe(imgs[0], poses[0], kf["kf_px"], kf["kf_level"], kf["pt_world"], kf["seed_px"], kf["seed_level"])
            s.set_last(imgs[0])
        cat = lambda key: np.concatenate([q[3][key] for q in seqs])
        trk.set_keyframe(np.stack([q[2][0] for q in seqs]), np.stack([q[1][0] for q in seqs]), np.arange(batch + 1) * N,
                         cat("kf_px"), cat("kf_level"), cat("pt_world"), np.arange(batch + 1) * S, cat("seed_px"), cat("seed_level"))
        trk.set_last(np.stack([q[2][0] for q in seqs]))
        n_flip = 0
        for k in range(1, n_frames + 1):
            stats, px, ok = trk.step(np.stack([q[2][k] for q in seqs]), np.stack([q[1][k - 1] for q in seqs]),
                                     np.concatenate([q[4][k - 1] for q in seqs]), want_px=True)
            seeds_g = trk.seeds()
            for b, s in enumerate(oseqs):
                e, pxe, oke = s.step(seqs[b][2][k], seqs[b][1][k - 1], seqs[b][4][k - 1], want_px=True)
                g = stats[b]
                assert g["n_tracked"] == e.n_tracked
                assert g["align_iters"] == e.align_iters, "GN iteration count differs (decision flip)"
                rot, trans = synth.pose_error(g["T_cur_w"], np.array(e.T_cur_w[:]))
                assert rot <= 1e-4 and trans <= 2e-4
                assert rot < 1e-9 and trans < 1e-9
                assert g["n_matched"] == e.n_matched and np.array_equal(ok[b * N:(b + 1) * N], oke)
                assert np.abs(px[b * N:(b + 1) * N] - pxe).max() <= 1e-3
                for key in ("n_seeds_updated", "n_seeds_converged", "n_seeds_failed", "n_seeds_skipped"):
                    if g[key] != getattr(e, key):
                        n_flip += 1
                se = s.seeds()
                sg = seeds_g[b * S:(b + 1) * S]
                sg = np.stack([sg[c] for c in ("a", "b", "mu", "z_range", "sigma2")], 1)
                close = np.isclose(sg, se, rtol=1e-5, atol=0).all(axis=1)
                assert close.mean() > 0.99, "seed states differ for %d of %d seeds" % ((~close).sum(), S)
                assert np.allclose(sg, se, rtol=1e-3, atol=0)
        assert n_flip == 0, "%d seed status count mismatches" % n_flip
    finally:
        trk.close()
        for s in oseqs:
            s.close()


@pytest.mark.gpu
def test_tracker_step_full_size_c4(ctx, oracle):
    """BASELINE.json configs[3] at its full size: one 1920x1080 sequence, 5-level pyramid, 1,000 features, 10,000 seeds,
    two frames against the oracle (sparse align runs on a thread-block cluster at this batch size)."""
    from android_svo_b200 import capi
    cfg, poses, imgs, kf, last_px = make_sequence(oracle, "C4", 0x00C0FFEE + 9, 3, tex_size=2048, allow_recycle=True)
    cam_o, cam_g = scenes.cam_of(cfg, Cam), scenes.cam_of(cfg, capi.Camera)
    args = (cfg["n_levels"], cfg["max_level"], cfg["min_level"], cfg["n_pyr"])
    N, S = cfg["n_features"], cfg["n_seeds"]
    assert (N, S) == (1000, 10000)
    so = OracleSeq(oracle, cam_o, *args)
    trk = capi.Tracker(ctx, cam_g, 1, *args)
    try:
        so.set_keyframe(imgs[0], poses[0], kf["kf_px"], kf["kf_level"], kf["pt_world"], kf["seed_px"], kf["seed_level"]); so.set_last(imgs[0])
        trk.set_keyframe(imgs[0][None], poses[0][None], [0, N], kf["kf_px"], kf["kf_level"], kf["pt_world"], [0, S], kf["seed_px"], kf["seed_level"])
        trk.set_last(imgs[0][None])
        for k in (1, 2):
            stats, px, ok = trk.step(imgs[k][None], poses[k - 1][None], last_px[k - 1], want_px=True)
            e, pxe, oke = so.step(imgs[k], poses[k - 1], last_px[k - 1], want_px=True)
            g = stats[0]
            assert g["n_tracked"] == e.n_tracked and g["align_iters"] == e.align_iters
            rot, trans = synth.pose_error(g["T_cur_w"], np.array(e.T_cur_w[:]))
            assert rot < 1e-9 and trans < 1e-9
            assert g["n_matched"] == e.n_matched and np.array_equal(ok, oke) and np.abs(px - pxe).max() <= 1e-3
            for key in ("n_seeds_updated", "n_seeds_converged", "n_seeds_failed", "n_seeds_skipped"):
                assert g[key] == getattr(e, key), key
            sg = trk.seeds()
            sg = np.stack([sg[c] for c in ("a", "b", "mu", "z_range", "sigma2")], 1)
            close = np.isclose(sg, so.seeds(), rtol=1e-5, atol=0).all(axis=1)
            assert close.mean() > 0.99
    finally:
        trk.close(); so.close()


@pytest.mark.gpu
def test_tracker_c5_size_replicas(ctx, oracle):
    """BASELINE.json configs[4] at its full size (4,096 sequences in one batch) through a size-independent property:
    the batch holds 32 distinct sequences, each replicated 128 times at scattered batch positions; every replica must
    return the bit-identical step record, refined pixels and seed state (no cross-talk, no dependence on the position
    in the batch or on which SM / group processed it), and replica 0 of each sequence must match the oracle."""
    from android_svo_b200 import capi
    n_distinct, B = 32, 4096
    seqs = [make_sequence(oracle, "C2", 0x00C0FFEE + 100 + i, 3) for i in range(n_distinct)]
    cfg = seqs[0][0]
    cam_o, cam_g = scenes.cam_of(cfg, Cam), scenes.cam_of(cfg, capi.Camera)
    args = (cfg["n_levels"], cfg["max_level"], cfg["min_level"], cfg["n_pyr"])
    N, S = cfg["n_features"], cfg["n_seeds"]
    which = (np.arange(B) * 13) % n_distinct                      # scattered assignment of sequences to batch slots
    trk = capi.Tracker(ctx, cam_g, B, *args)
    try:
        take = lambda key: np.concatenate([seqs[w][3][key] for w in which])
        trk.set_keyframe(np.stack([seqs[w][2][0] for w in which]), np.stack([seqs[w][1][0] for w in which]), np.arange(B + 1) * N,
                         take("kf_px"), take("kf_level"), take("pt_world"), np.arange(B + 1) * S, take("seed_px"), take("seed_level"))
        trk.set_last(np.stack([seqs[w][2][0] for w in which]))
        for k in (1, 2):
            stats, px, ok = trk.step(np.stack([seqs[w][2][k] for w in which]), np.stack([seqs[w][1][k - 1] for w in which]),
                                     np.concatenate([seqs[w][4][k - 1] for w in which]), want_px=True)
        seeds = trk.seeds()
        px = px.reshape(B, N, 2); ok = ok.reshape(B, N)
        seeds = np.stack([seeds[c] for c in ("a", "b", "mu", "z_range", "sigma2")], 1).reshape(B, S, 5)
        for d in range(n_distinct):
            idx = np.nonzero(which == d)[0]
            first = idx[0]
            for name in stats.dtype.names:
                assert (stats[name][idx] == stats[name][first]).all(), "replicas of sequence %d differ in %s" % (d, name)
            assert (px[idx] == px[first]).all() and (ok[idx] == ok[first]).all()
            assert (seeds[idx].view(np.uint32) == seeds[first].view(np.uint32)).all()
        for d in range(0, n_distinct, 8):                          # a few against the oracle
            so = OracleSeq(oracle, cam_o, *args)
            _, poses, imgs, kf, last_px = seqs[d]
            so.set_keyframe(imgs[0], poses[0], kf["kf_px"], kf["kf_level"], kf["pt_world"], kf["seed_px"], kf["seed_level"]); so.set_last(imgs[0])
            for k in (1, 2):
                e, pxe, oke = so.step(imgs[k], poses[k - 1], last_px[k - 1], want_px=True)
            first = np.nonzero(which == d)[0][0]
            g = stats[first]
            rot, trans = synth.pose_error(g["T_cur_w"], np.array(e.T_cur_w[:]))
            assert rot < 1e-9 and trans < 1e-9 and g["n_matched"] == e.n_matched and g["n_seeds_updated"] == e.n_seeds_updated
            assert np.array_equal(ok[first], oke) and np.abs(px[first] - pxe).max() <= 1e-3
            so.close()
    finally:
        trk.close()


def test_oracle_chain_matches_reference(oracle, ref):
    """Chain mode: Reprojector::reprojectMap + pose_optimizer::optimizeGaussNewton between alignment and the depth filter,
    as FrameHandlerMono::processFrame runs them (frame_handler_mono.cpp:191-222) — the oracle's restatement against the
    reference's own classes over a sequence (per-point reprojection counters reset every frame on both sides)."""
    cfg, poses, imgs, kf, last_px = make_sequence(oracle, "C2", 0x00C0FFEE + 3, 8)
    cam = scenes.cam_of(cfg, Cam)
    args = (cam, cfg["n_levels"], cfg["max_level"], cfg["min_level"], cfg["n_pyr"])
    so, sr = OracleSeq(oracle, *args), RefSeq(ref, *args)
    try:
        for s in (so, sr):
            s.set_keyframe(imgs[0], poses[0], kf["kf_px"], kf["kf_level"], kf["pt_world"], kf["seed_px"], kf["seed_level"])
            s.set_chain(30, 120, 1)
            s.set_last(imgs[0])
        for k in range(1, 8):
            a, pxa, oka = so.step(imgs[k], poses[k - 1], last_px[k - 1], want_px=True)
            b, pxb, okb = sr.step(imgs[k], poses[k - 1], last_px[k - 1], want_px=True)
            assert a.n_tracked == b.n_tracked and a.align_iters == b.align_iters
            assert (a.n_matched, a.n_reproj_trials, a.n_pose_obs) == (b.n_matched, b.n_reproj_trials, b.n_pose_obs)
            assert np.array_equal(oka, okb) and np.array_equal(pxa[oka == 1], pxb[okb == 1])     # the frame's new features: bit-exact
            assert np.array_equal(np.array(a.T_cur_w[:]), np.array(b.T_cur_w[:]))                # pose after the optimiser: bit-exact
            assert a.n_seeds_converged == b.n_seeds_converged
            assert np.isclose(so.seeds(), sr.seeds(), rtol=1e-6, atol=0).all(axis=1).mean() > 0.99
            # the pose optimiser tightens the alignment pose (which stops at level 2) against the ground truth
            grot, gtrans = synth.pose_error(np.array(a.T_cur_w[:]), poses[k])
            assert grot < 2e-3 and gtrans < 6e-3 and 40 < a.n_matched <= 121
    finally:
        so.close(); sr.close()


@pytest.mark.gpu
@pytest.mark.parametrize("pose_opt", [1, 0])
def test_tracker_chain_matches_oracle(ctx, oracle, pose_opt):
    """svob200_tracker_step in chain mode (reprojector grid rules + pose optimiser inside the step) against the oracle's
    chain, which is bit-exact to the reference's (test_oracle_chain_matches_reference)."""
    from android_svo_b200 import capi
    batch, n_frames = 3, 5
    seqs = [make_sequence(oracle, "C2", 0x00C0FFEE + 20 + i, n_frames + 1) for i in range(batch)]
    cfg = seqs[0][0]
    cam_o, cam_g = scenes.cam_of(cfg, Cam), scenes.cam_of(cfg, capi.Camera)
    args = (cfg["n_levels"], cfg["max_level"], cfg["min_level"], cfg["n_pyr"])
    oseqs = [OracleSeq(oracle, cam_o, *args) for _ in range(batch)]
    trk = capi.Tracker(ctx, cam_g, batch, *args)
    try:
        N, S = cfg["n_features"], cfg["n_seeds"]
        for s, (_, poses, imgs, kf, _) in zip(oseqs, seqs):
            s.set_keyframe(imgs[0], poses[0], kf["kf_px"], kf["kf_level"], kf["pt_world"], kf["seed_px"], kf["seed_level"])
            s.set_chain(30, 40, pose_opt)                       # max_fts = 40: the break fires
            s.set_last(imgs[0])
        cat = lambda key: np.concatenate([q[3][key] for q in seqs])
        trk.set_keyframe(np.stack([q[2][0] for q in seqs]), np.stack([q[1][0] for q in seqs]), np.arange(batch + 1) * N,
                         cat("kf_px"), cat("kf_level"), cat("pt_world"), np.arange(batch + 1) * S, cat("seed_px"), cat("seed_level"))
        trk.set_chain(30, 40, pose_opt)
        trk.set_last(np.stack([q[2][0] for q in seqs]))
        for k in range(1, n_frames + 1):
            stats, px, ok = trk.step(np.stack([q[2][k] for q in seqs]), np.stack([q[1][k - 1] for q in seqs]),
                                     np.concatenate([q[4][k - 1] for q in seqs]), want_px=True)
            seeds_g = trk.seeds()
            for b, s in enumerate(oseqs):
                e, pxe, oke = s.step(seqs[b][2][k], seqs[b][1][k - 1], seqs[b][4][k - 1], want_px=True)
                g = stats[b]
                assert g["n_tracked"] == e.n_tracked and g["align_iters"] == e.align_iters
                assert (g["n_matched"], g["n_reproj_trials"], g["n_pose_obs"]) == (e.n_matched, e.n_reproj_trials, e.n_pose_obs)
                assert g["n_matched"] == 41
                okb = ok[b * N:(b + 1) * N]
                assert np.array_equal(okb, oke) and np.abs(px[b * N:(b + 1) * N][okb == 1] - pxe[oke == 1]).max() <= 1e-3
                rot, trans = synth.pose_error(g["T_cur_w"], np.array(e.T_cur_w[:]))
                assert rot < 1e-9 and trans < 1e-9
                for key in ("n_seeds_updated", "n_seeds_converged", "n_seeds_failed", "n_seeds_skipped"):
                    assert g[key] == getattr(e, key), key
                sg = seeds_g[b * S:(b + 1) * S]
                sg = np.stack([sg[c] for c in ("a", "b", "mu", "z_range", "sigma2")], 1)
                assert np.isclose(sg, s.seeds(), rtol=1e-5, atol=0).all(axis=1).mean() > 0.99
    finally:
        trk.close()
        for s in oseqs:
            s.close()
