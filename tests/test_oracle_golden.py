"""The CPU oracle (oracle/svo_oracle.c) against the committed golden vectors that were generated
from the REAL reference compiled for the host and from python cv2 (tests/golden/make_golden.py).
Runs everywhere (no GPU, no /root/reference)."""
import os
import numpy as np
import pytest

from android_svo_b200 import synth, frontend
from oracle.pyoracle import Cam, Seed, OracleSeq

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))


def cam_of_golden():
    c = G["scene_cam"]
    return Cam.make(int(c[0]), int(c[1]), *c[2:]), dict(w=int(c[0]), h=int(c[1]), fx=c[2], fy=c[3], cx=c[4], cy=c[5])


def test_fast_matches_cv2(oracle):
    for i in range(4):
        img = G["fast_img%d" % i]
        xs, ys, sc = oracle.fast(img, 10, True)
        assert np.array_equal(np.stack([xs, ys, sc], 1), G["fast_nms%d" % i])
        xs, ys, _ = oracle.fast(img, 10, False)
        assert np.array_equal(np.stack([xs, ys], 1), G["fast_raw%d" % i])


def test_pyramid_both_roundings(oracle):
    img = G["pyr_img"]
    a = oracle.pyramid(img, 4)                     # x86 dispatch: 160, 80 -> SSE2; 40 -> scalar (40 % 16 != 0)
    b = oracle.pyramid(img, 4, modes=[0, 0, 0])
    for l in range(1, 4):
        assert np.array_equal(a[l], G["pyr_sse2_l%d" % l])
        assert np.array_equal(b[l], G["pyr_trunc_l%d" % l])
    assert np.array_equal(oracle.half_sample(G["pyr_odd_img"], 0), G["pyr_odd_out"])     # odd-width scalar walk


def test_shi_tomasi_and_zmssd(oracle):
    img, img0 = G["st_img"], G["pyr_img"]
    got = np.array([oracle.shi_tomasi(img, u, v) for u, v in G["st_uv"]], np.float32)
    assert np.array_equal(got.view(np.uint32), G["st_score"].view(np.uint32))
    got = [oracle.zmssd(img[y0 - 4:y0 + 4, x0 - 4:x0 + 4].copy().reshape(-1), img0, x1, y1) for x0, y0, x1, y1 in G["zm_xy"]]
    assert np.array_equal(np.array(got, np.int32), G["zm_score"])


def test_fast_detect(oracle):
    cam, cfg = cam_of_golden()
    pyr = oracle.pyramid(G["scene_imgs"][0], 4)
    n, cells = oracle.fast_detect(pyr, 4, 20, 10.0)
    sel = cells[cells["score"] > 10.0]
    assert np.array_equal(np.stack([sel["x"], sel["y"]], 1).astype(np.float64), G["detect_px"])
    assert np.array_equal(sel["level"], G["detect_level"])


def test_se3_exp(oracle):
    for v, e in zip(G["se3_exp_in"], G["se3_exp_out"]):
        assert oracle.se3_exp(v).tobytes() == e.tobytes()


def test_sparse_align(oracle):
    cam, cfg = cam_of_golden()
    imgs, poses = G["scene_imgs"], G["scene_poses"]
    px, ptw = G["align_px"], G["align_ptw"]
    has = (~np.isnan(ptw[:, 0])).astype(np.uint8)
    # derived inputs exactly as the reference computes them
    f = np.array([oracle.cam2world(cam, p[0], p[1]) for p in px])
    assert f.reshape(-1).tobytes() == G["align_f"].tobytes()
    n, res = oracle.sparse_align(oracle.pyramid(imgs[0], 4), oracle.pyramid(imgs[1], 4), cam, px.reshape(-1), G["align_xyz_ref"], has,
                                 G["align_T_cur_ref_init"], 3, 1)
    chi2, n_meas, stop, n_tracked = G["align_scalars"]
    assert n == n_tracked and res.n_meas == n_meas and res.stop == stop
    assert list(res.iters) == list(G["align_iters"])
    assert res.chi2 == chi2
    assert np.array(res.H[:]).tobytes() == G["align_H"].tobytes()
    T_cur_w = oracle.se3_mul(np.array(res.T_cur_ref[:]), poses[0])
    rot, trans = synth.pose_error(T_cur_w, G["align_T_cur_w"])
    assert rot < 1e-12 and trans < 1e-12


def test_align2d_align1d(oracle):
    tgt = G["scene_imgs"][2]
    for i in range(len(G["lk_px0"])):
        ok, p = oracle.align2d(tgt, G["lk_pwb"][i], G["lk_patch"][i], 10, G["lk_px0"][i])
        assert ok == G["lk_ok2"][i] and p.tobytes() == G["lk_px2"][i].tobytes()
        ok, p, hi = oracle.align1d(tgt, G["lk_dir"][i], G["lk_pwb"][i], G["lk_patch"][i], 10, G["lk_px0"][i])
        assert ok == G["lk_ok1"][i] and p.tobytes() == G["lk_px1"][i].tobytes() and hi == G["lk_hinv"][i]


def test_epipolar_match(oracle):
    cam, cfg = cam_of_golden()
    imgs, poses = G["scene_imgs"], G["scene_poses"]
    pyr = [oracle.pyramid(im, 4) for im in imgs]
    opts = oracle.matcher_opts(4)
    n_ok = 0
    for i in range(len(G["epi_px"])):
        c = int(G["epi_cur"][i])
        ftr = oracle.ref_feature(G["epi_px"][i], G["epi_f"][i], int(G["epi_level"][i]))
        T_cur_ref = oracle.se3_mul(poses[c], oracle.se3_inverse(poses[0]))
        d = G["epi_d"][i]
        ok, e = oracle.find_epipolar_match(pyr[0], pyr[c], cam, ftr, T_cur_ref, d[0], d[1], d[2], opts)
        assert ok == G["epi_ok"][i]
        assert np.array(e.A_cur_ref[:]).tobytes() == G["epi_A"][i].tobytes()
        assert e.search_level == G["epi_sl"][i] and e.epi_length == G["epi_epi_len"][i]
        assert np.array_equal(np.array(e.patch_with_border[:], np.uint8), G["epi_pwb"][i])
        if ok:
            assert e.depth == G["epi_depth"][i] and np.array(e.px_cur[:]).tobytes() == G["epi_px_cur"][i].tobytes()
            n_ok += 1
    assert n_ok > 50


def test_update_seed_and_tau(oracle):
    for i in range(len(G["seed_x"])):
        s = Seed(*[float(v) for v in G["seed_in"][i]])
        oracle.update_seed(float(G["seed_x"][i]), float(G["seed_tau2"][i]), s)
        got = np.array([s.a, s.b, s.mu, s.z_range, s.sigma2], np.float32)
        assert got.tobytes() == G["seed_out"][i].tobytes()
        assert oracle.compute_tau(G["tau_T"][i], G["tau_f"][i], float(G["tau_z"][i]), float(G["tau_ang"])) == G["tau_out"][i]


def test_frontend_step_pipeline(oracle):
    cam, cfg = cam_of_golden()
    imgs, poses = G["scene_imgs"], G["scene_poses"]
    seq = OracleSeq(oracle, cam, 4, 3, 1, 4)
    try:
        seq.set_keyframe(imgs[0], poses[0], G["pipe_kf_px"], G["pipe_kf_level"], G["pipe_pt_world"], G["pipe_seed_px"], G["pipe_seed_level"])
        seq.set_last(imgs[0])
        for k in range(1, 6):
            s, px, ok = seq.step(imgs[k], poses[k - 1], G["pipe_last_px"][k - 1], want_px=True)
            assert np.array_equal(np.array(s.T_cur_w[:]), G["pipe_T"][k - 1]), "pose not bit-identical to the reference's"
            assert [s.n_tracked, s.n_matched, s.n_seeds_converged, s.align_iters] == list(G["pipe_counts"][k - 1])
            assert np.array_equal(ok, G["pipe_ok"][k - 1]) and np.array_equal(px, G["pipe_px"][k - 1])
            # the restatement reproduces the reference's seed states bit for bit
            assert np.array_equal(seq.seeds().view(np.uint32), np.ascontiguousarray(G["pipe_seeds"][k - 1], dtype=np.float32).view(np.uint32))
    finally:
        seq.close()
