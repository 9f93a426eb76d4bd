"""GPU parity of the callers either side of the hot path through the C ABI: reprojector, pose optimizer,
point optimizer, YUV->gray input stage fused with the pyramid, seed initialisation — against the CPU oracle
(oracle/svo_oracle_map.c, itself pinned to the real reference in tests/test_map_oracle.py) and against the committed
golden vectors generated from the reference (tests/golden/map_golden.npz)."""
import os
import numpy as np
import pytest

from android_svo_b200 import capi, synth
from oracle.pyoracle import Cam
from oracle import pyoracle_map as pm
import map_scenes as ms

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "map_golden.npz")


@pytest.fixture(scope="module")
def om(oracle):
    return pm.OracleMap(oracle)


def ocam(cfg):
    return Cam.make(cfg["w"], cfg["h"], cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"])


def gcam(cfg):
    return capi.Camera.make(cfg["w"], cfg["h"], cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"])


def to_gpu_records(pts, obs, T_kf, kf_fids, image, obs_base=0):
    """oracle-side arrays -> the ABI's svob200_map_point / svob200_feature_ref (+ T_obs_w)"""
    gp = np.zeros(len(pts), capi.map_point_dt)
    gp["pos"] = pts["pos"]; gp["type"] = pts["type"]; gp["obs_begin"] = pts["obs_begin"] + obs_base; gp["obs_end"] = pts["obs_end"] + obs_base
    go = np.zeros(len(obs), capi.feature_ref_dt)
    go["ref_frame_id"] = [kf_fids[k] for k in obs["keyframe"]]
    go["ref_image"] = image
    go["level"] = obs["ftr"]["level_ref"]; go["type"] = obs["ftr"]["type"]; go["px"] = obs["ftr"]["px_ref"]
    go["f"] = obs["ftr"]["f_ref"]; go["grad"] = obs["ftr"]["grad"]
    return gp, go, T_kf[obs["keyframe"]]


def run_oracle(om, oracle, sc, max_fts):
    cfg = sc["cfg"]
    order = ms.insertion_order(sc)
    pts, obs = ms.reorder(sc, order)
    kf_pyrs = [oracle.pyramid(i, cfg["n_levels"]) for i in sc["kf_imgs"]]
    cur = oracle.pyramid(sc["cur_img"], cfg["n_levels"])
    res, winner, nm, nt = om.reproject_map(kf_pyrs, cur, ocam(cfg), sc["T_cur"], pts, obs, sc["T_kf"], sc["cell"], max_fts, oracle.matcher_opts(cfg["n_pyr"]))
    return order, pts, obs, res, winner, nm, nt


def check_reproj(got, want, base=0):
    res, winner, stats = got
    ores, owinner, nm, nt = want
    assert np.array_equal(res["status"], ores["status"])
    assert np.array_equal(res["cell"], ores["cell"])
    tried = np.isin(ores["status"], (pm.REPROJ_FAILED, pm.REPROJ_MATCHED))
    assert np.array_equal(res["obs"][tried] - base, ores["obs"][tried])
    assert np.array_equal(res["search_level"][tried], ores["search_level"][tried])
    assert np.array_equal(res["px"], ores["px"])                       # bit-exact: projections, refined and failed positions
    assert np.array_equal(res["A_cur_ref"][tried], ores["A_cur_ref"][tried])
    assert np.array_equal(winner, owinner)
    assert (stats["n_matches"], stats["n_trials"]) == (nm, nt)


@pytest.mark.parametrize("seed,max_fts", [(5, 120), (6, 120), (7, 12), (8, 0)])
def test_reprojector_vs_oracle(ctx, om, oracle, seed, max_fts):
    sc = ms.build_map_scene(oracle, seed=seed)
    cfg = sc["cfg"]
    order, pts, obs, ores, owinner, nm, nt = run_oracle(om, oracle, sc, max_fts)
    kf_fids = [100 + k for k in range(len(sc["kf_imgs"]))]
    for fid, img in zip(kf_fids, sc["kf_imgs"]):
        ctx.frame_create(fid, 1, cfg["w"], cfg["h"], cfg["n_levels"]); ctx.frame_upload(fid, img)
    ctx.frame_create(99, 1, cfg["w"], cfg["h"], cfg["n_levels"]); ctx.frame_upload(99, sc["cur_img"])
    gp, go, To = to_gpu_records(pts, obs, sc["T_kf"], kf_fids, 0)
    res, winner, stats = ctx.reproject_map(99, gcam(cfg), sc["T_cur"], [0, len(gp)], gp, go, To, sc["cell"], max_fts, ctx.matcher_opts(cfg["n_pyr"]))
    check_reproj((res, winner[0], stats[0]), (ores, owinner, nm, nt))
    assert stats[0]["n_in_frame"] == int((ores["status"] != pm.REPROJ_NOT_IN_FRAME).sum())
    for fid in kf_fids + [99]:
        ctx.frame_release(fid)


def test_reprojector_batch_and_golden(ctx, om, oracle):
    """two independent maps in one call (image 0 / image 1 of every frame batch) + the reference's own output"""
    scs = [ms.build_map_scene(oracle, seed=s) for s in (5, 9)]
    cfg = scs[0]["cfg"]
    n_kf = len(scs[0]["kf_imgs"])
    kf_fids = [200 + k for k in range(n_kf)]
    for k, fid in enumerate(kf_fids):
        ctx.frame_create(fid, 2, cfg["w"], cfg["h"], cfg["n_levels"]); ctx.frame_upload(fid, np.stack([sc["kf_imgs"][k] for sc in scs]))
    ctx.frame_create(199, 2, cfg["w"], cfg["h"], cfg["n_levels"]); ctx.frame_upload(199, np.stack([sc["cur_img"] for sc in scs]))
    want, gps, gos, Tos, orders = [], [], [], [], []
    obs_base, off = 0, [0]
    for b, sc in enumerate(scs):
        order, pts, obs, ores, owinner, nm, nt = run_oracle(om, oracle, sc, 120)
        gp, go, To = to_gpu_records(pts, obs, sc["T_kf"], kf_fids, b, obs_base)
        want.append((ores, owinner, nm, nt, obs_base)); gps.append(gp); gos.append(go); Tos.append(To); orders.append(order)
        obs_base += len(obs); off.append(off[-1] + len(gp))
    res, winner, stats = ctx.reproject_map(199, gcam(cfg), np.stack([sc["T_cur"] for sc in scs]), off, np.concatenate(gps), np.concatenate(gos),
                                           np.concatenate(Tos), scs[0]["cell"], 120, ctx.matcher_opts(cfg["n_pyr"]))
    for b in range(2):
        ores, owinner, nm, nt, base = want[b]
        w = winner[b].copy(); w[w >= 0] -= off[b]
        check_reproj((res[off[b]:off[b + 1]], w, stats[b]), (ores, owinner, nm, nt), base)
    # image 0 is the golden scene: new features of the frame = winners in cell order, pixels bit-exact with the reference
    G = np.load(GOLD)
    assert int(G["reproj_seed"]) == 5
    w0 = winner[0][winner[0] >= 0]
    assert np.array_equal(orders[0][w0], G["reproj_new_point"])
    assert np.array_equal(res["px"][w0], G["reproj_new_px"]) and np.array_equal(res["search_level"][w0], G["reproj_new_level"])
    assert stats[0]["n_matches"] == int(G["reproj_n_matches"]) and stats[0]["n_trials"] == int(G["reproj_n_trials"])
    for fid in kf_fids + [199]:
        ctx.frame_release(fid)


def test_pose_optimizer_vs_oracle_and_golden(ctx, om):
    """batch of 5 problems incl. an empty one; sums are tree-reduced on the device => pose within 1e-9 rad / 1e-9 m of the
    sequential reference (tolerance stated; the spec allows 1e-4), discrete outputs equal"""
    scenes = [ms.pose_opt_scene(seed=s) for s in (3, 4, 9, 12)]
    cfg = scenes[0]["cfg"]
    cam_o, cam_g = ocam(cfg), gcam(cfg)
    fs = [np.array([om.o.cam2world(cam_o, p[0], p[1]) for p in s["px"]]) for s in scenes]
    off = [0]
    for s in scenes[:2]:
        off.append(off[-1] + len(s["level"]))
    off.append(off[-1])                                 # an image without features: returns untouched (errors.empty())
    for s in scenes[2:]:
        off.append(off[-1] + len(s["level"]))
    T0 = np.stack([scenes[0]["T_init"], scenes[1]["T_init"], scenes[0]["T_init"], scenes[2]["T_init"], scenes[3]["T_init"]])
    T, res, outl = ctx.pose_optimize(cam_g, off, np.concatenate(fs), np.concatenate([s["level"] for s in scenes]),
                                     np.concatenate([s["pos"] for s in scenes]), T0)
    assert np.array_equal(T[2], T0[2]) and res[2]["num_obs"] == 0 and res[2]["iters"] == 0
    G = np.load(GOLD)
    for b, k in ((0, 0), (1, 1), (3, 2), (4, 3)):
        s = scenes[k]
        To, ro, oo = om.pose_optimize(cam_o, fs[k], s["level"], s["pos"], s["T_init"])
        rot, trans = synth.pose_error(T[b], To)
        assert rot <= 1e-9 and trans <= 1e-9
        assert np.array_equal(outl[off[b]:off[b + 1]], oo) and res[b]["num_obs"] == ro["num_obs"]
        assert res[b]["iters"] == ro["iters"] and res[b]["rolled_back"] == ro["rolled_back"]
        assert res[b]["estimated_scale"] == ro["estimated_scale"] and res[b]["error_init"] == ro["error_init"]     # order statistics: exact
        assert np.isclose(res[b]["error_final"], ro["error_final"], rtol=1e-9)
        assert np.allclose(res[b]["A"], ro["A"], rtol=1e-9, atol=1e-9 * np.abs(ro["A"]).max())
        assert np.isclose(res[b]["chi2"], ro["chi2"], rtol=1e-9)
    rot, trans = synth.pose_error(T[0], G["pose_T"])     # the reference's own result for scene 0
    assert rot <= 1e-9 and trans <= 1e-9 and np.array_equal(outl[off[0]:off[1]], G["pose_outlier"])


def test_point_optimizer_vs_oracle_and_golden(ctx, om):
    pts = ms.point_opt_scene()
    off = np.concatenate([[0], np.cumsum([len(p["T"]) for p in pts])])
    got, iters = ctx.points_optimize(off, np.concatenate([p["T"] for p in pts]), np.concatenate([p["f"] for p in pts]),
                                     np.stack([p["pos0"] for p in pts]))
    G = np.load(GOLD)
    for i, p in enumerate(pts):
        want, it = om.point_optimize(p["T"], p["f"], p["pos0"])
        assert np.array_equal(got[i], want) and iters[i] == it          # bit-exact: same operation order, no libm on the path
    assert np.array_equal(got, G["point_pos"])


@pytest.mark.parametrize("w,h,ps,pad,u_first,levels", [(64, 48, 2, 0, False, 3), (64, 48, 2, 0, True, 3), (64, 48, 1, 0, False, 3),
                                                       (640, 480, 2, 0, False, 4), (752, 480, 2, 16, False, 5), (200, 120, 1, 8, False, 3),
                                                       (94, 60, 2, 2, False, 2), (46, 31, 2, 0, False, 1)])
def test_yuv_input_stage(ctx, om, oracle, w, h, ps, pad, u_first, levels):
    """YUV_420_888 planes -> gray level 0 -> pyramid, fused: every level bit-exact with the oracle's YUV2RGB + RGBA2GRAY +
    halfSample; planar and interleaved chroma, padded strides, sizes the fused kernel takes and sizes it does not"""
    frames = [ms.yuv_frame(w, h, 10 + b, ps, pad) for b in range(2)]
    if u_first and ps == 2:
        for fr in frames:
            fr["u"], fr["v"] = fr["v"], fr["u"]         # NV12: U is the lower address
    fid = 300
    ctx.frame_create(fid, 2, w, h, levels)
    # batch layout: image b's planes sit at fixed strides inside one buffer each
    ybuf = np.stack([fr["y"] for fr in frames])
    y_img = ybuf[0].nbytes
    fr0 = frames[0]
    if ps == 2:
        n = len(fr0["u"]) + 1
        cbuf = np.zeros((2, n), np.uint8)
        for b, fr in enumerate(frames):
            lo, hi = (fr["v"], fr["u"]) if not u_first else (fr["u"], fr["v"])
            cbuf[b, :-1] = lo; cbuf[b, -1] = hi[-1]
        lo_view, hi_view = cbuf.reshape(-1)[0:], cbuf.reshape(-1)[1:]
        u_arr, v_arr = (hi_view, lo_view) if not u_first else (lo_view, hi_view)
        ctx.frame_upload_yuv420(fid, ybuf, u_arr, v_arr, fr0["y_stride"], fr0["uv_stride"], 2, y_img, n)
    else:
        ubuf = np.stack([fr["u"] for fr in frames]); vbuf = np.stack([fr["v"] for fr in frames])
        ctx.frame_upload_yuv420(fid, ybuf, ubuf, vbuf, fr0["y_stride"], fr0["uv_stride"], 1, y_img, ubuf[0].nbytes)
    for b, fr in enumerate(frames):
        gray = om.yuv420_to_gray(fr["y"], fr["u"], fr["v"], fr["uv_stride"], fr["uv_pixel_stride"], w, h, fr["y_stride"])
        pyr = oracle.pyramid(gray, levels)
        for l in range(levels):
            assert np.array_equal(ctx.frame_download(fid, b, l), pyr[l]), "level %d of image %d differs" % (l, b)
    ctx.frame_release(fid)


def test_yuv_golden(ctx):
    G = np.load(GOLD)
    fr = ms.yuv_frame(64, 48, 1, 2, 0)
    ctx.frame_create(301, 1, 64, 48, 2)
    ctx.frame_upload_yuv420(301, fr["y"], fr["u"], fr["v"], fr["y_stride"], fr["uv_stride"], 2)
    assert np.array_equal(ctx.frame_download(301, 0, 0), G["yuv_gray"])     # cv2.cvtColor of the app's RGBA
    ctx.frame_release(301)


def test_seed_init_vs_oracle_and_golden(ctx, om, oracle):
    cfg = ms.SMALL
    tex = synth.make_texture(512)
    imgs = [synth.render(tex, cfg, synth.trajectory(4, seed=2)[k]) for k in (2, 3)]
    rng = np.random.RandomState(0)
    ex = [np.c_[rng.uniform(0, cfg["w"], 40), rng.uniform(0, cfg["h"], 40)], np.zeros((0, 2))]
    ex[1] = np.c_[rng.uniform(0, cfg["w"], 15), rng.uniform(0, cfg["h"], 15)]
    dm, dn = [2.2, 3.1], [1.7, 0.9]
    ctx.frame_create(310, 2, cfg["w"], cfg["h"], cfg["n_levels"]); ctx.frame_upload(310, np.stack(imgs))
    corners, seeds, counts = ctx.seeds_initialize(310, cfg["n_pyr"], 20, 8.0, [0, 40, 55], np.concatenate(ex), dm, dn)
    G = np.load(GOLD)
    for b in range(2):
        oc, osd = om.initialize_seeds(oracle.pyramid(imgs[b], cfg["n_levels"]), ocam(cfg), cfg["n_pyr"], 20, 8.0, ex[b], dm[b], dn[b])
        n = counts[b]
        assert n == len(oc) > 20
        for k in ("x", "y", "level"):
            assert np.array_equal(corners[b][:n][k], oc[k])
        assert np.array_equal(corners[b][:n]["score"].view(np.uint32), oc["score"].view(np.uint32))
        for k in ("a", "b", "mu", "z_range", "sigma2"):
            assert np.array_equal(seeds[b][:n][k].view(np.uint32), osd[k].view(np.uint32))
    n = counts[0]
    assert np.array_equal(np.stack([corners[0][:n][k] for k in ("x", "y", "level")], 1), G["seedinit_xyl"])
    assert np.array_equal(np.stack([seeds[0][:n][k] for k in ("a", "b", "mu", "z_range", "sigma2")], 1).view(np.uint32), G["seedinit_seeds"].view(np.uint32))
    ctx.frame_release(310)
