"""Deterministic map-side scenes (keyframes, map points with observation lists, point candidates) for the
reprojector / pose-optimizer / point-optimizer / seed-init tests.  Used by the CPU tests (oracle vs the real
reference), the GPU tests (CUDA vs oracle / golden) and tests/golden/make_golden_map.py."""
import numpy as np

from android_svo_b200 import synth
from oracle import pyoracle_map as pm

SMALL = dict(w=320, h=240, fx=262.5, fy=262.5, cx=159.5, cy=119.5, n_levels=4, n_pyr=4)


def project(cfg, T, p):
    pc = synth.se3_transform(T, p)
    return np.array([cfg["fx"] * pc[0] / pc[2] + cfg["cx"], cfg["fy"] * pc[1] / pc[2] + cfg["cy"]]), pc[2]


def bearing(cfg, px):
    v = np.array([(px[0] - cfg["cx"]) / cfg["fx"], (px[1] - cfg["cy"]) / cfg["fy"], 1.0])
    return v / np.linalg.norm(v)


def build_map_scene(oracle, cfg=SMALL, seed=5, n_kf=3, cell=30, tex_size=512, n_candidates=12, with_edgelets=True, bad_frac=0.2):
    """Keyframes at trajectory poses 0, 10, 20 ...; the current frame a little further along.  Map points come from
    the FAST corners of EVERY keyframe (so grid cells hold several candidates), observed by their own keyframe and,
    for about half of them, by a second keyframe.  A fraction gets a wrong 3D position (their alignment fails), a
    few are TYPE_DELETED, the last `n_candidates` points are point candidates (one observation each)."""
    rng = np.random.RandomState(seed)
    tex = synth.make_texture(tex_size)
    poses = synth.trajectory(10 * n_kf + 8, seed=seed + 1, amp_scale=1.2)
    T_kf = np.array([poses[10 * k] for k in range(n_kf)])
    T_cur = poses[10 * (n_kf - 1) + 5].copy()
    kf_imgs = [synth.render(tex, cfg, T) for T in T_kf]
    cur_img = synth.render(tex, cfg, T_cur)
    pos, typ, obs_of = [], [], []
    for k in range(n_kf):
        pyr = oracle.pyramid(kf_imgs[k], cfg["n_levels"])
        _, cells = oracle.fast_detect(pyr, cfg["n_pyr"], cell, 12.0)
        good = cells[cells["score"].astype(np.float64) > 12.0]
        for c in good[:: 2 if k else 1]:
            px = np.array([float(c["x"]), float(c["y"])])
            p = synth.backproject_to_plane(cfg, T_kf[k], px)
            obs = [(k, px, int(c["level"]))]
            k2 = (k + 1 + rng.randint(n_kf - 1)) % n_kf if n_kf > 1 else k
            if k2 != k and rng.rand() < 0.5:
                px2, z2 = project(cfg, T_kf[k2], p)
                if 12 <= px2[0] < cfg["w"] - 12 and 12 <= px2[1] < cfg["h"] - 12:
                    # the second observation goes FIRST in about half of the cases (list order matters on cos ties only,
                    # but it exercises the arg-max over the list)
                    o2 = (k2, px2, int(rng.randint(0, 2)))
                    obs = [o2] + obs if rng.rand() < 0.5 else obs + [o2]
            if rng.rand() < bad_frac:
                p = p + np.array([rng.uniform(-0.12, 0.12), rng.uniform(-0.12, 0.12), rng.uniform(-0.3, 0.3)])
            pos.append(p)
            r = rng.rand()
            typ.append(pm.POINT_GOOD if r < 0.3 else (pm.POINT_DELETED if r > 0.93 else pm.POINT_UNKNOWN))
            obs_of.append(obs)
    n_map = len(pos)
    # candidates: extra corners of the last keyframe on a finer grid
    pyr = oracle.pyramid(kf_imgs[-1], cfg["n_levels"])
    _, cells = oracle.fast_detect(pyr, cfg["n_pyr"], 16, 10.0)
    good = cells[cells["score"].astype(np.float64) > 10.0]
    for c in good[3::max(1, len(good) // max(n_candidates, 1))][:n_candidates]:
        px = np.array([float(c["x"]), float(c["y"])])
        pos.append(synth.backproject_to_plane(cfg, T_kf[-1], px))
        typ.append(pm.POINT_CANDIDATE)
        obs_of.append([(n_kf - 1, px, int(c["level"]))])
    n = len(pos)
    points = np.zeros(n, pm.map_point_dt)
    obs = np.zeros(sum(len(o) for o in obs_of), pm.point_obs_dt)
    j = 0
    for i in range(n):
        points[i]["pos"] = pos[i]; points[i]["type"] = typ[i]; points[i]["obs_begin"] = j
        for (k, px, lv) in obs_of[i]:
            obs[j]["keyframe"] = k
            obs[j]["ftr"]["px_ref"] = px; obs[j]["ftr"]["f_ref"] = bearing(cfg, px); obs[j]["ftr"]["level_ref"] = lv
            is_edge = with_edgelets and rng.rand() < 0.15
            obs[j]["ftr"]["type"] = 1 if is_edge else 0
            a = rng.uniform(0, 2 * np.pi)
            obs[j]["ftr"]["grad"] = (np.cos(a), np.sin(a)) if is_edge else (1.0, 0.0)
            j += 1
        points[i]["obs_end"] = j
    return dict(cfg=cfg, T_kf=T_kf, T_cur=T_cur, kf_imgs=kf_imgs, cur_img=cur_img, points=points, obs=obs, n_map=n_map,
                n_candidates=n - n_map, cell=cell)


def insertion_order(sc):
    """Order in which Reprojector::reprojectMap puts the points into its grid (reprojector.cpp:94-145) for a map built
    the way ref_harness_map.cpp builds it: keyframes sorted by |t_cur - t_kf| of T_f_w (map.cpp:126), each keyframe's
    fts_ in point-index order, every point once; then the candidates."""
    T_kf, T_cur, points, obs, n_map = sc["T_kf"], sc["T_cur"], sc["points"], sc["obs"], sc["n_map"]
    d = np.linalg.norm(T_kf[:, :3] - T_cur[None, :3], axis=1)
    seen, order = set(), []
    for k in np.argsort(d, kind="stable"):
        for i in range(n_map):
            if i in seen:
                continue
            if any(obs[j]["keyframe"] == k for j in range(points[i]["obs_begin"], points[i]["obs_end"])):
                seen.add(i); order.append(i)
    order += list(range(n_map, len(points)))
    return np.array(order, np.int64)


def reorder(sc, order):
    """points/obs permuted into insertion order (obs ranges rebuilt)."""
    pts = sc["points"][order].copy()
    new_obs = np.zeros(len(sc["obs"]), pm.point_obs_dt)
    j = 0
    for i, src in enumerate(order):
        b, e = sc["points"][src]["obs_begin"], sc["points"][src]["obs_end"]
        new_obs[j:j + e - b] = sc["obs"][b:e]
        pts[i]["obs_begin"] = j; pts[i]["obs_end"] = j + e - b
        j += e - b
    return pts, new_obs


def pose_opt_scene(cfg=SMALL, seed=3, n=90, noise_px=0.4, outlier_frac=0.08):
    """n features of one frame with their 3D points; the frame pose is perturbed; a few gross outliers."""
    rng = np.random.RandomState(seed)
    T_true = synth.trajectory(30, seed=seed)[17]
    px = np.c_[rng.uniform(12, cfg["w"] - 12, n), rng.uniform(12, cfg["h"] - 12, n)]
    pos = np.array([synth.backproject_to_plane(cfg, T_true, p) for p in px])
    pos[:, 2] += rng.uniform(-0.3, 0.3, n)           # not a plane
    px = np.array([project(cfg, T_true, p)[0] for p in pos])
    px += rng.normal(0, noise_px, px.shape)
    out = rng.rand(n) < outlier_frac
    px[out] += rng.uniform(-25, 25, (int(out.sum()), 2))
    level = rng.randint(0, 3, n).astype(np.int32)
    d = synth.se3_from_rotvec_trans(rng.uniform(-0.01, 0.01, 3), rng.uniform(-0.02, 0.02, 3))
    T_init = synth.se3_mul(d, T_true)
    f = np.array([bearing(cfg, p) for p in px])
    return dict(cfg=cfg, px=px, f=f, level=level, pos=pos, T_init=T_init, T_true=T_true)


def point_opt_scene(cfg=SMALL, seed=4, n_points=40, max_obs=5):
    """points observed from 2..max_obs keyframes with pixel noise, starting from a perturbed position."""
    rng = np.random.RandomState(seed)
    poses = synth.trajectory(60, seed=seed, amp_scale=2.0)
    out = []
    for i in range(n_points):
        p = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.4, 0.4), rng.uniform(1.6, 2.6)])
        ks = rng.choice(60, rng.randint(2, max_obs + 1), replace=False)
        T = poses[ks]
        f = []
        for t in T:
            px, _ = project(cfg, t, p)
            f.append(bearing(cfg, px + rng.normal(0, 0.3, 2)))
        out.append(dict(T=np.array(T), f=np.array(f), pos0=p + rng.normal(0, 0.03, 3), pos_true=p))
    return out


def yuv_frame(w, h, seed, uv_pixel_stride=2, pad=0):
    """A synthetic YUV_420_888 image the way AImage hands it out: Y plane (stride w+pad) and U/V planes that are either
    planar (pixel stride 1) or views into one interleaved VU buffer (pixel stride 2, NV21-like)."""
    rng = np.random.RandomState(seed)
    ys = w + pad
    y = rng.randint(0, 256, (h, ys)).astype(np.uint8)
    # smooth-ish luma so the gray image is not pure noise, but keep full range incl. < 16 and > 235
    y[:, :w] = np.clip((np.add.outer(np.arange(h) * 255 // max(h - 1, 1), np.arange(w) * 64 // max(w - 1, 1)) // 1 + rng.randint(-40, 40, (h, w))), 0, 255)
    cw, ch = (w + 1) // 2, (h + 1) // 2
    if uv_pixel_stride == 2:
        uvs = 2 * cw + pad
        buf = rng.randint(0, 256, ch * uvs + 1).astype(np.uint8)
        v = buf[:-1]          # V first (NV21): v[o], u = v + 1
        u = buf[1:]
    else:
        uvs = cw + pad
        u = rng.randint(0, 256, ch * uvs).astype(np.uint8)
        v = rng.randint(0, 256, ch * uvs).astype(np.uint8)
    return dict(y=y, y_stride=ys, u=u, v=v, uv_stride=uvs, uv_pixel_stride=uv_pixel_stride, w=w, h=h)
