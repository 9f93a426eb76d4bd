"""Scene set-up for the per-frame front-end step (host side, numpy): which corners become map
features / seeds, their ground-truth 3D points, and the per-frame inputs (pose of the last frame,
pixel of every map point in the last frame).  Follows SURVEY.md §8(d).  No compute path here — the
step itself is svob200_tracker_step (csrc/tracker.cu).
"""
import numpy as np
from . import synth


def select_features(cells, thr, n, recycle=False):
    """First n cells (cell-index order, like FastDetector::detect's output order) whose score > thr.
    recycle: a view with fewer corners repeats the ones it has, so that every sequence carries the same load (bench only)."""
    good = cells[cells["score"].astype(np.float64) > thr]
    if len(good) < n and recycle and len(good) > 0:
        good = np.resize(good, n)
    if len(good) < n:
        raise ValueError("only %d corners above %.1f, need %d" % (len(good), thr, n))
    good = good[:n]
    return np.stack([good["x"], good["y"]], 1).astype(np.float64), good["level"].astype(np.int32)


def project(cfg, T_f_w, pts):
    """Level-0 pixels of world points seen from pose T_f_w (numpy; input generation only)."""
    out = np.zeros((len(pts), 2))
    for i, p in enumerate(pts):
        pc = synth.se3_transform(T_f_w, p)
        out[i] = [cfg["fx"] * pc[0] / pc[2] + cfg["cx"], cfg["fy"] * pc[1] / pc[2] + cfg["cy"]]
    return out


def project_many(cfg, T_f_w, pts):
    """Vectorised project()."""
    q = np.asarray(T_f_w[3:7]); t = np.asarray(T_f_w[:3])
    qv = q[:3]
    uv = np.cross(qv, pts)
    uv = uv + uv
    pc = pts + q[3] * uv + np.cross(qv, uv) + t
    return np.stack([cfg["fx"] * pc[:, 0] / pc[:, 2] + cfg["cx"], cfg["fy"] * pc[:, 1] / pc[:, 2] + cfg["cy"]], 1)


def keyframe_setup(cfg, T_kf_w, ftr_cells, seed_cells, ftr_thr, seed_thr, plane_z=2.0, recycle=False):
    """Map features + seeds of one keyframe from two detector passes (coarse grid / fine grid)."""
    kf_px, kf_level = select_features(ftr_cells, ftr_thr, cfg["n_features"], recycle)
    seed_px, seed_level = select_features(seed_cells, seed_thr, cfg["n_seeds"], recycle)
    pt_world = np.array([synth.backproject_to_plane(cfg, T_kf_w, p, plane_z) for p in kf_px])
    return dict(kf_px=kf_px, kf_level=kf_level, pt_world=pt_world, seed_px=seed_px, seed_level=seed_level)


# detector grids per config: (feature cell, feature thr, seed cell, seed thr)
DETECT = {"C2": (40, 20.0, 20, 10.0), "C3": (30, 20.0, 12, 10.0), "C4": (40, 20.0, 14, 10.0)}
DETECT["C1"] = DETECT["C5"] = DETECT["C2"]
