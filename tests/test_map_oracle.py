"""CPU: the C restatement of the callers either side of the hot path (oracle/svo_oracle_map.c) against the REAL
reference compiled for this host (oracle/_ref/libsvo_ref.so via ref_harness_map.cpp) and against the committed
golden vectors generated from it (tests/golden/map_golden.npz, make_golden_map.py)."""
import os
import numpy as np
import pytest

from android_svo_b200 import synth
from oracle.pyoracle import Cam, Pyramid
from oracle import pyoracle_map as pm
import map_scenes as ms

GOLD = os.path.join(os.path.dirname(__file__), "golden", "map_golden.npz")


@pytest.fixture(scope="module")
def om(oracle):
    return pm.OracleMap(oracle)


@pytest.fixture(scope="module")
def rm(ref):
    r = pm.RefMap(ref)
    if not r.available():
        pytest.skip("oracle/_ref/libsvo_ref.so lacks the map harness")
    return r


def cam_of(cfg):
    return Cam.make(cfg["w"], cfg["h"], cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"])


def oracle_reproject(om, oracle, sc, max_fts):
    cfg = sc["cfg"]
    cam = cam_of(cfg)
    order = ms.insertion_order(sc)
    pts, obs = ms.reorder(sc, order)
    kf_pyrs = [oracle.pyramid(i, cfg["n_levels"]) for i in sc["kf_imgs"]]
    cur = oracle.pyramid(sc["cur_img"], cfg["n_levels"])
    mo = oracle.matcher_opts(cfg["n_pyr"])
    res, winner, nm, nt = om.reproject_map(kf_pyrs, cur, cam, sc["T_cur"], pts, obs, sc["T_kf"], sc["cell"], max_fts, mo)
    return order, pts, obs, res, winner, nm, nt


def expected_from_status(sc, order, pts, obs, res, winner):
    """what the reference's side effects must be, read off the per-point status"""
    n = len(pts)
    failed = np.zeros(n, np.int32); succ = np.zeros(n, np.int32)
    for i in range(n):
        src = order[i]
        if res[i]["status"] == pm.REPROJ_FAILED:
            failed[src] += 1
        if res[i]["status"] == pm.REPROJ_MATCHED:
            succ[src] += 1
        if src >= sc["n_map"] and res[i]["status"] == pm.REPROJ_NOT_IN_FRAME:
            failed[src] += 3
    new = []
    for c in range(len(winner)):
        i = winner[c]
        if i < 0:
            continue
        o = obs[res[i]["obs"]]
        g = np.array([0.0, 0.0])          # Feature ctor default grad is (1,0) for corners
        if o["ftr"]["type"] == 1:
            A = res[i]["A_cur_ref"].reshape(2, 2)
            g = A @ o["ftr"]["grad"]
            g = g / np.linalg.norm(g)
        new.append((order[i], res[i]["px"][0], res[i]["px"][1], res[i]["search_level"], int(o["ftr"]["type"]), g[0], g[1]))
    return failed, succ, np.array(new).reshape(-1, 7)


@pytest.mark.parametrize("seed,max_fts", [(5, 120), (6, 120), (7, 12)])
def test_reprojector_matches_reference(om, oracle, rm, seed, max_fts):
    sc = ms.build_map_scene(oracle, seed=seed)
    cfg = sc["cfg"]
    rm.config(cfg["n_pyr"], sc["cell"], max_fts)
    r = rm.reproject_map(sc["kf_imgs"], sc["T_kf"], sc["cur_img"], sc["T_cur"], cam_of(cfg), sc["points"], sc["obs"], sc["n_candidates"])
    order, pts, obs, res, winner, nm, nt = oracle_reproject(om, oracle, sc, max_fts)
    assert (nm, nt) == (r["n_matches"], r["n_trials"])
    failed, succ, new = expected_from_status(sc, order, pts, obs, res, winner)
    assert np.array_equal(failed, r["n_failed"]) and np.array_equal(succ, r["n_succeeded"])
    assert len(new) == len(r["new_point"]) == nm
    assert np.array_equal(new[:, 0].astype(np.int64), r["new_point"])
    assert np.array_equal(new[:, 1:3], r["new_px"])                      # bit-exact pixel positions
    assert np.array_equal(new[:, 3].astype(np.int32), r["new_level"]) and np.array_equal(new[:, 4].astype(np.int32), r["new_type"])
    edge = new[:, 4] == 1
    assert np.allclose(new[edge][:, 5:7], r["new_grad"][edge], rtol=0, atol=1e-12)
    # the scene exercises every branch
    st = res["status"]
    if max_fts > 100:
        for s in (pm.REPROJ_NOT_IN_FRAME, pm.REPROJ_UNTRIED, pm.REPROJ_FAILED, pm.REPROJ_MATCHED):
            assert (st == s).any(), "status %d not exercised" % s
    else:
        assert nm == max_fts + 1                                          # the maxFts break (:164) fired


def test_pose_optimizer_matches_reference(om, rm):
    for seed in (3, 4, 9):
        s = ms.pose_opt_scene(seed=seed)
        cam = cam_of(s["cfg"])
        img = np.zeros((s["cfg"]["h"], s["cfg"]["w"]), np.uint8)
        r = rm.pose_optimize(cam, img, s["px"], s["level"], s["pos"], s["T_init"])
        # the reference recomputes f = cam2world(px) in the Feature ctor: feed the oracle the same bearing
        f = np.array([om.o.cam2world(cam, p[0], p[1]) for p in s["px"]])
        T, res, outl = om.pose_optimize(cam, f, s["level"], s["pos"], s["T_init"])
        assert np.array_equal(T, r["T"])                                  # bit-exact (same libm, Eigen's LDLT association restated)
        assert np.array_equal(outl, r["outlier"]) and res["num_obs"] == r["num_obs"]
        assert np.allclose(res["A"], r["A"], rtol=1e-6, atol=1e-6 * np.abs(r["A"]).max())
        assert np.isclose(res["estimated_scale"], r["estimated_scale"], rtol=1e-6)
        assert np.isclose(res["error_init"], r["error_init"], rtol=1e-9) and np.isclose(res["error_final"], r["error_final"], rtol=1e-6)
        assert outl.sum() > 0


def test_point_optimizer_matches_reference(om, rm):
    cfg = ms.SMALL
    cam = cam_of(cfg)
    img = np.zeros((cfg["h"], cfg["w"]), np.uint8)
    before, after = [], []
    for p in ms.point_opt_scene():
        want = rm.point_optimize(cam, img, p["T"], p["f"], p["pos0"])
        got, _ = om.point_optimize(p["T"], p["f"], p["pos0"])
        assert np.array_equal(got, want)                                  # bit-exact
        before.append(np.linalg.norm(p["pos0"] - p["pos_true"])); after.append(np.linalg.norm(got - p["pos_true"]))
    assert np.median(after) < np.median(before)                            # the scene is a meaningful optimisation problem


def test_seed_init_matches_reference(om, oracle, rm):
    cfg = ms.SMALL
    cam = cam_of(cfg)
    tex = synth.make_texture(512)
    img = synth.render(tex, cfg, synth.trajectory(4, seed=2)[2])
    rng = np.random.RandomState(0)
    existing = np.c_[rng.uniform(0, cfg["w"], 40), rng.uniform(0, cfg["h"], 40)]
    rm.config(cfg["n_pyr"], 30, 120)
    xs, ys, lv, seeds = rm.initialize_seeds(cam, img, cfg["n_pyr"], 20, 8.0, existing, 2.2, 1.7)
    corners, oseeds = om.initialize_seeds(oracle.pyramid(img, cfg["n_levels"]), cam, cfg["n_pyr"], 20, 8.0, existing, 2.2, 1.7)
    assert len(xs) == len(corners) > 20
    assert np.array_equal(corners["x"], xs) and np.array_equal(corners["y"], ys) and np.array_equal(corners["level"], lv)
    got = np.stack([oseeds[k] for k in ("a", "b", "mu", "z_range", "sigma2")], 1)
    assert np.array_equal(got.view(np.uint32), seeds.view(np.uint32))


def test_yuv_to_gray_matches_cv2(om):
    """YUV2RGB is the app's own integer code (image_process.cpp:97-126); RGBA2GRAY is third-party (cv::cvtColor), pinned
    against python cv2 when it is importable and against the golden fixture otherwise."""
    cv2 = pytest.importorskip("cv2")
    for seed, ps, pad in ((1, 2, 0), (2, 1, 0), (3, 2, 16)):
        fr = ms.yuv_frame(64, 48, seed, ps, pad)
        rgba = om.yuv420_to_rgba(fr["y"], fr["u"], fr["v"], fr["uv_stride"], fr["uv_pixel_stride"], fr["w"], fr["h"], fr["y_stride"])
        assert np.array_equal(om.rgba_to_gray(rgba), cv2.cvtColor(rgba, cv2.COLOR_RGBA2GRAY))
        assert np.array_equal(om.yuv420_to_gray(fr["y"], fr["u"], fr["v"], fr["uv_stride"], fr["uv_pixel_stride"], fr["w"], fr["h"], fr["y_stride"]),
                              cv2.cvtColor(rgba, cv2.COLOR_RGBA2GRAY))
    # exhaustive over the channel cube edges: every (c0, c1, c2) with two channels on a coarse lattice
    v = np.arange(256, dtype=np.uint8)
    cube = np.stack(np.meshgrid(v, v[::5], v[::7], indexing="ij"), -1).reshape(1, -1, 3)
    rgba = np.concatenate([cube, np.full(cube.shape[:2] + (1,), 255, np.uint8)], -1)
    assert np.array_equal(om.rgba_to_gray(rgba), cv2.cvtColor(rgba, cv2.COLOR_RGBA2GRAY))


@pytest.mark.skipif(not os.path.exists(GOLD), reason="tests/golden/map_golden.npz not generated")
def test_oracle_against_map_golden(om, oracle):
    G = np.load(GOLD)
    # reprojector
    sc = ms.build_map_scene(oracle, seed=int(G["reproj_seed"]))
    order, pts, obs, res, winner, nm, nt = oracle_reproject(om, oracle, sc, int(G["reproj_max_fts"]))
    failed, succ, new = expected_from_status(sc, order, pts, obs, res, winner)
    assert (nm, nt) == (int(G["reproj_n_matches"]), int(G["reproj_n_trials"]))
    assert np.array_equal(failed, G["reproj_n_failed"]) and np.array_equal(succ, G["reproj_n_succeeded"])
    assert np.array_equal(new[:, 0].astype(np.int64), G["reproj_new_point"]) and np.array_equal(new[:, 1:3], G["reproj_new_px"])
    # pose optimizer
    s = ms.pose_opt_scene(seed=int(G["pose_seed"]))
    cam = cam_of(s["cfg"])
    f = np.array([om.o.cam2world(cam, p[0], p[1]) for p in s["px"]])
    T, r, outl = om.pose_optimize(cam, f, s["level"], s["pos"], s["T_init"])
    assert np.array_equal(T, G["pose_T"]) and np.array_equal(outl, G["pose_outlier"])
    # point optimizer
    for i, p in enumerate(ms.point_opt_scene()):
        got, _ = om.point_optimize(p["T"], p["f"], p["pos0"])
        assert np.array_equal(got, G["point_pos"][i])
    # input stage
    fr = ms.yuv_frame(64, 48, 1, 2, 0)
    assert np.array_equal(om.yuv420_to_gray(fr["y"], fr["u"], fr["v"], fr["uv_stride"], fr["uv_pixel_stride"], fr["w"], fr["h"], fr["y_stride"]),
                          G["yuv_gray"])
    # seed init
    cfg = ms.SMALL
    img = synth.render(synth.make_texture(512), cfg, synth.trajectory(4, seed=2)[2])
    rng = np.random.RandomState(0)
    existing = np.c_[rng.uniform(0, cfg["w"], 40), rng.uniform(0, cfg["h"], 40)]
    corners, oseeds = om.initialize_seeds(oracle.pyramid(img, cfg["n_levels"]), cam_of(cfg), cfg["n_pyr"], 20, 8.0, existing, 2.2, 1.7)
    assert np.array_equal(np.stack([corners["x"], corners["y"], corners["level"]], 1), G["seedinit_xyl"])
    assert np.array_equal(np.stack([oseeds[k] for k in ("a", "b", "mu", "z_range", "sigma2")], 1).view(np.uint32), G["seedinit_seeds"].view(np.uint32))
