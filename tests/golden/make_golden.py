"""Generates the committed golden vectors under tests/golden/ by running
  * the REAL reference (oracle/_ref/libsvo_ref.so: the reference's own sources compiled for this host,
    x86 SSE2 paths; libsvo_ref_nosse.so for the truncating halfSample) and
  * python cv2 (cv::FAST, the one third-party algorithm on the path)
on small seeded inputs.  Needs /root/reference (to build oracle/_ref) and cv2; run from the repo root:

    python tests/golden/make_golden.py

The fixtures pin the C restatement (oracle/svo_oracle.c) and the CUDA path on machines where the
reference itself is not available.
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from android_svo_b200 import synth, frontend  # noqa: E402
from oracle.pyoracle import Oracle, Ref, Cam, RefSeq  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CFG = dict(w=320, h=240, fx=262.5, fy=262.5, cx=159.5, cy=119.5, n_levels=4, n_pyr=4, n_features=60, n_seeds=150,
           max_level=3, min_level=1)


def noise(h, w, seed, blur):
    rng = np.random.RandomState(seed)
    img = rng.randint(0, 256, (h, w)).astype(np.uint8)
    if blur:
        f = img.astype(np.int32)
        acc = np.zeros_like(f)
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                acc += np.roll(np.roll(f, dy, 0), dx, 1)
        img = (acc // 9).astype(np.uint8)
    return img


def main():
    import cv2
    ref, ref_ns, oracle = Ref(), Ref(nosse=True), Oracle()
    assert ref.available() and ref_ns.available(), "build oracle/_ref first (make -C oracle ref)"
    g = {}
    # ---------------- cv::FAST (python cv2)
    det = cv2.FastFeatureDetector_create(10, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    det0 = cv2.FastFeatureDetector_create(10, False, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    for i, (h, w) in enumerate([(60, 80), (67, 120), (30, 47), (96, 128)]):
        img = noise(h, w, 100 + i, True)
        g["fast_img%d" % i] = img
        k = det.detect(img, None)
        g["fast_nms%d" % i] = np.array([(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in k], np.int32).reshape(-1, 3)
        k = det0.detect(img, None)
        g["fast_raw%d" % i] = np.array([(int(p.pt[0]), int(p.pt[1])) for p in k], np.int32).reshape(-1, 2)
    g["cv2_version"] = np.array(cv2.__version__)
    # ---------------- pyramid (both roundings), Shi-Tomasi, ZMSSD
    img = noise(120, 160, 7, False)
    g["pyr_img"] = img
    for l, lv in enumerate(ref.pyramid(img, 4).levels[1:]):
        g["pyr_sse2_l%d" % (l + 1)] = lv
    for l, lv in enumerate(ref_ns.pyramid(img, 4).levels[1:]):
        g["pyr_trunc_l%d" % (l + 1)] = lv
    odd = noise(30, 47, 8, False)
    g["pyr_odd_img"] = odd
    g["pyr_odd_out"] = ref.half_sample(odd)
    img2 = noise(120, 160, 9, True)
    g["st_img"] = img2
    rng = np.random.RandomState(1)
    uv = np.stack([rng.randint(0, 160, 300), rng.randint(0, 120, 300)], 1)
    g["st_uv"] = uv.astype(np.int32)
    g["st_score"] = np.array([ref.shi_tomasi(img2, u, v) for u, v in uv], np.float32)
    xy = np.stack([rng.randint(4, 156, 200), rng.randint(4, 116, 200), rng.randint(4, 156, 200), rng.randint(4, 116, 200)], 1)
    g["zm_xy"] = xy.astype(np.int32)
    g["zm_score"] = np.array([ref.zmssd(img2[y0 - 4:y0 + 4, x0 - 4:x0 + 4].copy().reshape(-1), img, x1, y1) for x0, y0, x1, y1 in xy], np.int32)
    # ---------------- scene for the geometric operators
    tex = synth.make_texture(512)
    poses = synth.trajectory(40, seed=11, amp_scale=1.5)[::4]
    imgs = [synth.render(tex, CFG, T) for T in poses[:6]]
    cam = Cam.make(CFG["w"], CFG["h"], CFG["fx"], CFG["fy"], CFG["cx"], CFG["cy"])
    g["scene_imgs"] = np.stack(imgs)
    g["scene_poses"] = poses[:6]
    g["scene_cam"] = np.array([CFG["w"], CFG["h"], CFG["fx"], CFG["fy"], CFG["cx"], CFG["cy"]])
    ref.config(CFG["n_pyr"], CFG["n_levels"] - 1, CFG["min_level"])
    # FastDetector::detect
    px, lv = ref.fast_detect(imgs[0], cam, CFG["n_pyr"], 20, 10.0)
    g["detect_px"], g["detect_level"] = px, lv
    # SparseImgAlign::run
    N = 80
    apx = np.c_[rng.uniform(10, 310, N), rng.uniform(10, 230, N)]
    apx[:4] = [[2, 2], [318, 238], [24, 24], [160, 3]]
    ptw = np.array([synth.backproject_to_plane(CFG, poses[0], p) for p in apx])
    ptw[5] = np.nan
    r = ref.sparse_align(imgs[0], imgs[1], cam, CFG["max_level"], CFG["min_level"], 30, poses[0], poses[0], apx.reshape(-1), ptw.reshape(-1))
    g["align_px"], g["align_ptw"] = apx, ptw
    for k in ("T_cur_w", "T_cur_ref_init", "H", "Jres", "x", "f", "xyz_ref", "iters"):
        g["align_" + k] = r[k]
    g["align_scalars"] = np.array([r["chi2"], r["n_meas"], r["stop"], r["n_tracked"]])
    # align2D / align1D
    src, tgt = imgs[0], imgs[2]
    n = 120
    A2 = dict(pwb=[], patch=[], px0=[], dir=[], ok2=[], px2=[], ok1=[], px1=[], hinv=[])
    for i in range(n):
        x, y = rng.randint(6, 314), rng.randint(6, 234)
        pwb = src[y - 5:y + 5, x - 5:x + 5].copy().reshape(-1)
        patch = src[y - 4:y + 4, x - 4:x + 4].copy().reshape(-1)
        p0 = np.array([x + rng.uniform(-2, 2), y + rng.uniform(-2, 2)])
        d = rng.randn(2).astype(np.float32); d /= np.linalg.norm(d)
        ok2, p2 = ref.align2d(tgt, pwb, patch, 10, p0)
        ok1, p1, hi = ref.align1d(tgt, d, pwb, patch, 10, p0)
        for k, v in zip(A2.keys(), (pwb, patch, p0, d, ok2, p2, ok1, p1, hi)):
            A2[k].append(v)
    for k, v in A2.items():
        g["lk_" + k] = np.array(v)
    # Matcher::findEpipolarMatchDirect
    n = 150
    E = dict(px=[], level=[], d=[], cur=[], ok=[], depth=[], px_cur=[], epi_len=[], sl=[], A=[], f=[], pwb=[])
    for i in range(n):
        lvl = rng.randint(0, 3)
        p = np.floor(np.array([rng.uniform(12, 308), rng.uniform(12, 228)]) / (1 << lvl)) * (1 << lvl)
        zgt = np.linalg.norm(synth.backproject_to_plane(CFG, poses[0], p) - synth.se3_inverse(poses[0])[:3])
        mu = 1.0 / (zgt * rng.uniform(0.7, 1.4)); sig = rng.choice([0.3, 0.1, 0.02, 0.002])
        d = [1 / mu, 1 / (mu + sig), 1 / max(mu - sig, 1e-8)]
        c = rng.randint(1, 6)
        r = ref.find_epipolar_match(imgs[0], imgs[c], cam, poses[0], poses[c], p, lvl, *d)
        for k, v in zip(E.keys(), (p, lvl, d, c, r["success"], r["depth"], r["px_cur"], r["epi_length"], r["search_level"], r["A"], r["f_ref"], r["pwb"])):
            E[k].append(v)
    for k, v in E.items():
        g["epi_" + k] = np.array(v)
    # DepthFilter::updateSeed / computeTau
    n = 300
    st = np.stack([rng.uniform(5, 30, n), rng.uniform(5, 30, n), rng.uniform(0.2, 1.0, n), rng.uniform(0.5, 2.0, n), rng.uniform(1e-4, 0.1, n)], 1).astype(np.float32)
    x = rng.uniform(0.2, 1.0, n).astype(np.float32); tau2 = rng.uniform(1e-6, 1e-2, n).astype(np.float32)
    g["seed_in"], g["seed_x"], g["seed_tau2"] = st, x, tau2
    g["seed_out"] = np.array([ref.update_seed(float(x[i]), float(tau2[i]), st[i]) for i in range(n)], np.float32)
    T = np.array([oracle.se3_exp(rng.randn(6) * [0.2, 0.2, 0.05, 0.02, 0.02, 0.02]) for _ in range(n)])
    f = np.array([ref.cam2world(cam, rng.uniform(0, 320), rng.uniform(0, 240)) for _ in range(n)])
    z = rng.uniform(1, 4, n); ang = 2 * np.arctan(1 / (2 * CFG["fx"]))
    g["tau_T"], g["tau_f"], g["tau_z"], g["tau_ang"] = T, f, z, np.array(ang)
    g["tau_out"] = np.array([ref.compute_tau(T[i], f[i], z[i], ang) for i in range(n)])
    g["se3_exp_in"] = rng.randn(50, 6) * [0.1, 0.1, 0.1, 0.05, 0.05, 0.05]
    g["se3_exp_out"] = np.array([ref.se3_exp(v) for v in g["se3_exp_in"]])
    # ---------------- the per-frame step through the reference classes (pipeline golden)
    pyr = oracle.pyramid(imgs[0], CFG["n_levels"])
    _, fcells = oracle.fast_detect(pyr, CFG["n_pyr"], 20, 12.0)
    _, scells = oracle.fast_detect(pyr, CFG["n_pyr"], 10, 8.0)
    kf = frontend.keyframe_setup(CFG, poses[0], fcells, scells, 12.0, 8.0)
    seq = RefSeq(ref, cam, CFG["n_levels"], CFG["max_level"], CFG["min_level"], CFG["n_pyr"])
    seq.set_keyframe(imgs[0], poses[0], kf["kf_px"], kf["kf_level"], kf["pt_world"], kf["seed_px"], kf["seed_level"])
    seq.set_last(imgs[0])
    for k, v in kf.items():
        g["pipe_" + k] = v
    P = dict(T=[], counts=[], px=[], ok=[], seeds=[], last_px=[])
    for k in range(1, 6):
        lp = frontend.project_many(CFG, poses[k - 1], kf["pt_world"])
        s, px, ok = seq.step(imgs[k], poses[k - 1], lp, want_px=True)
        for key, v in zip(P.keys(), (np.array(s.T_cur_w[:]), [s.n_tracked, s.n_matched, s.n_seeds_converged, s.align_iters], px, ok, seq.seeds(), lp)):
            P[key].append(v)
    seq.close()
    for k, v in P.items():
        g["pipe_" + k] = np.array(v)
    path = os.path.join(OUT, "reference_golden.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(g), "arrays")


if __name__ == "__main__":
    main()
