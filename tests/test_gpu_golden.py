"""CUDA path (through the C ABI) against the committed golden vectors produced by the REAL reference
(tests/golden/make_golden.py) — the parity check that does not go through the C restatement."""
import os
import numpy as np
import pytest

from android_svo_b200 import capi, synth

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))
_id = [5000]


def up(ctx, imgs, n_levels, modes=None):
    _id[0] += 1
    arr = np.stack(imgs)
    ctx.frame_create(_id[0], arr.shape[0], arr.shape[2], arr.shape[1], n_levels)
    ctx.frame_upload(_id[0], arr, round_modes=modes)
    return _id[0]


def cam_g():
    c = G["scene_cam"]
    return capi.Camera.make(int(c[0]), int(c[1]), *c[2:])


def test_fast_vs_cv2(ctx):
    for i in range(4):
        fid = up(ctx, [G["fast_img%d" % i]], 1)
        xs, ys, sc = ctx.fast_corners(fid, 0, 0, 10, True)
        assert np.array_equal(np.stack([xs, ys, sc], 1), G["fast_nms%d" % i])
        xs, ys, _ = ctx.fast_corners(fid, 0, 0, 10, False)
        assert np.array_equal(np.stack([xs, ys], 1), G["fast_raw%d" % i])
        ctx.frame_release(fid)


def test_pyramid_vs_reference(ctx):
    for modes, key in ((None, "sse2"), ([0, 0, 0], "trunc")):
        fid = up(ctx, [G["pyr_img"]], 4, modes)
        for l in range(1, 4):
            assert np.array_equal(ctx.frame_download(fid, 0, l), G["pyr_%s_l%d" % (key, l)])
        ctx.frame_release(fid)
    assert np.array_equal(ctx.half_sample(G["pyr_odd_img"], capi.ROUND_TRUNC), G["pyr_odd_out"])


def test_detect_vs_reference(ctx):
    fid = up(ctx, [G["scene_imgs"][0]], 4)
    cells, counts = ctx.fast_detect(fid, 4, 20, 10.0)
    sel = cells[0][cells[0]["score"] > 10.0]
    assert counts[0] == len(G["detect_px"])
    assert np.array_equal(np.stack([sel["x"], sel["y"]], 1).astype(np.float64), G["detect_px"])
    assert np.array_equal(sel["level"], G["detect_level"])
    ctx.frame_release(fid)


def test_sparse_align_vs_reference(ctx):
    imgs, poses = G["scene_imgs"], G["scene_poses"]
    rid, cid = up(ctx, [imgs[0]], 4), up(ctx, [imgs[1]], 4)
    px, ptw = G["align_px"], G["align_ptw"]
    has = (~np.isnan(ptw[:, 0])).astype(np.uint8)
    res = ctx.sparse_align(rid, cid, cam_g(), [0, len(px)], px, G["align_xyz_ref"].reshape(-1, 3), has, G["align_T_cur_ref_init"][None, :], 3, 1)[0]
    chi2, n_meas, stop, n_tracked = G["align_scalars"]
    assert res["n_meas"] == n_meas and list(res["iters"]) == list(G["align_iters"]) and res["stop"] == stop
    assert abs(res["chi2"] - chi2) <= 1e-5 * chi2
    # compose like SparseImgAlign::run does and compare the frame pose
    T = np.zeros((1, 7))
    ctx._ck(ctx.L.svob200_compose_poses(ctx.h, 1, capi._ptr(np.array([res])), capi._ptr(np.ascontiguousarray(poses[0])), capi._ptr(T), capi.MEM_HOST))
    rot, trans = synth.pose_error(T[0], G["align_T_cur_w"])
    assert rot <= 1e-4 and trans <= 2e-4 and rot < 1e-9 and trans < 1e-9
    assert np.abs(res["H"] - G["align_H"]).max() <= 1e-9 * np.abs(G["align_H"]).max()
    ctx.frame_release(rid); ctx.frame_release(cid)


def test_lk_vs_reference(ctx):
    fid = up(ctx, [G["scene_imgs"][2]], 1)
    n = len(G["lk_px0"])
    conv, px, _ = ctx.align_patches(fid, 0, np.zeros(n, np.int32), G["lk_pwb"], G["lk_patch"], 10, G["lk_px0"])
    assert np.array_equal(conv, G["lk_ok2"].astype(np.int32)) and px.tobytes() == G["lk_px2"].tobytes()
    conv, px, hi = ctx.align_patches(fid, 0, np.zeros(n, np.int32), G["lk_pwb"], G["lk_patch"], 10, G["lk_px0"], dirv=G["lk_dir"])
    assert np.array_equal(conv, G["lk_ok1"].astype(np.int32)) and px.tobytes() == G["lk_px1"].tobytes() and np.array_equal(hi, G["lk_hinv"])
    ctx.frame_release(fid)


def test_epipolar_vs_reference(ctx, oracle):
    imgs, poses = G["scene_imgs"], G["scene_poses"]
    rid = up(ctx, [imgs[0]], 4)
    cid = up(ctx, list(imgs), 4)            # all six views as one batch: cur_image selects the view
    n = len(G["epi_px"])
    f = capi.make_feature_refs(n)
    f["ref_frame_id"] = rid
    f["cur_image"] = G["epi_cur"]
    f["level"] = G["epi_level"]
    f["px"], f["f"] = G["epi_px"], G["epi_f"]
    f["grad"] = (1.0, 0.0)
    for i in range(n):
        f[i]["T_cur_ref"] = oracle.se3_mul(poses[int(G["epi_cur"][i])], oracle.se3_inverse(poses[0]))
    got = ctx.epipolar_match(cid, cam_g(), f, G["epi_d"], ctx.matcher_opts(4))
    assert np.array_equal(got["success"], G["epi_ok"].astype(np.int32))
    assert got["A_cur_ref"].tobytes() == G["epi_A"].tobytes()
    assert np.array_equal(got["search_level"], G["epi_sl"]) and np.array_equal(got["epi_length"], G["epi_epi_len"])
    assert np.array_equal(got["patch_with_border"], G["epi_pwb"]), "warped patches differ from the reference"
    ok = G["epi_ok"].astype(bool)
    assert got["px_cur"][ok].tobytes() == G["epi_px_cur"][ok].tobytes()
    assert np.abs(got["depth"][ok] - G["epi_depth"][ok]).max() <= 1e-12 * G["epi_depth"][ok].max()
    ctx.frame_release(rid); ctx.frame_release(cid)


def test_seed_scalars_vs_reference(ctx):
    seeds = np.zeros(len(G["seed_x"]), capi.seed_dt)
    for k, c in enumerate(("a", "b", "mu", "z_range", "sigma2")):
        seeds[c] = G["seed_in"][:, k]
    got = ctx.update_seed(G["seed_x"], G["seed_tau2"], seeds)
    out = np.stack([got[c] for c in ("a", "b", "mu", "z_range", "sigma2")], 1)
    assert np.allclose(out, G["seed_out"], rtol=1e-5, atol=0)
    assert (out.view(np.uint32) == G["seed_out"].view(np.uint32)).all(axis=1).mean() > 0.9
    tau = ctx.compute_tau(G["tau_T"], G["tau_f"], G["tau_z"], float(G["tau_ang"]))
    assert np.allclose(tau, G["tau_out"], rtol=1e-9, atol=0)


def test_tracker_vs_reference_pipeline(ctx, oracle):
    """svob200_tracker_step against the reference's own per-frame outputs (golden).  Seed states: the free-running oracle is
    first pinned to the golden seeds BIT FOR BIT (so it stands for the reference), then test_pipeline.SeedParity applies: seeds
    outside 1e-5 are counted and each must be reproduced exactly by the oracle run on the device's pose."""
    from oracle.pyoracle import Cam, OracleSeq
    from test_pipeline import SeedParity
    imgs, poses = G["scene_imgs"], G["scene_poses"]
    N, S = len(G["pipe_kf_level"]), len(G["pipe_seed_level"])
    cam = cam_g()
    cam_o = Cam.make(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy)
    trk = capi.Tracker(ctx, cam, 1, 4, 3, 1, 4)
    free, pinned = OracleSeq(oracle, cam_o, 4, 3, 1, 4), OracleSeq(oracle, cam_o, 4, 3, 1, 4)
    par = SeedParity()
    try:
        trk.set_keyframe(imgs[:1], poses[:1], [0, N], G["pipe_kf_px"], G["pipe_kf_level"], G["pipe_pt_world"], [0, S],
                         G["pipe_seed_px"], G["pipe_seed_level"])
        trk.set_last(imgs[:1])
        for s in (free, pinned):
            s.set_keyframe(imgs[0], poses[0], G["pipe_kf_px"], G["pipe_kf_level"], G["pipe_pt_world"], G["pipe_seed_px"], G["pipe_seed_level"])
            s.set_last(imgs[0])
        for k in range(1, 6):
            st, px, ok = trk.step(imgs[k:k + 1], poses[k - 1:k], G["pipe_last_px"][k - 1], want_px=True)
            rot, trans = synth.pose_error(st[0]["T_cur_w"], G["pipe_T"][k - 1])
            assert rot <= 1e-4 and trans <= 2e-4 and rot < 1e-9 and trans < 1e-9
            assert [st[0]["n_tracked"], st[0]["n_matched"], st[0]["n_seeds_converged"], st[0]["align_iters"]] == list(G["pipe_counts"][k - 1])
            assert np.array_equal(ok, G["pipe_ok"][k - 1]) and np.abs(px - G["pipe_px"][k - 1]).max() <= 1e-3
            free.step(imgs[k], poses[k - 1], G["pipe_last_px"][k - 1])
            assert np.array_equal(free.seeds().view(np.uint32), np.ascontiguousarray(G["pipe_seeds"][k - 1], dtype=np.float32).view(np.uint32))
            pinned.set_pose_override(st[0]["T_cur_w"])
            pinned.step(imgs[k], poses[k - 1], G["pipe_last_px"][k - 1])
            par.check(trk.seeds(), trk.seed_obs(), free, pinned)
        par.finish(max_outside_frac=0.02)
    finally:
        trk.close(); free.close(); pinned.close()
