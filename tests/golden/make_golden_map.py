"""Generates tests/golden/map_golden.npz: outputs of the REAL reference (oracle/_ref/libsvo_ref.so built from the
reference's own reprojector.cpp / map.cpp / pose_optimizer.cpp / point.cpp / depth_filter.cpp) and of python cv2
(cv::cvtColor RGBA2GRAY, third-party) on the seeded scenes of tests/map_scenes.py.  Needs /root/reference and cv2:

    python tests/golden/make_golden_map.py
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from android_svo_b200 import synth  # noqa: E402
from oracle.pyoracle import Oracle, Cam  # noqa: E402
from oracle import pyoracle_map as pm  # noqa: E402
import map_scenes as ms  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def cam_of(cfg):
    return Cam.make(cfg["w"], cfg["h"], cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"])


def main():
    import cv2
    oracle = Oracle()
    om, rm = pm.OracleMap(oracle), pm.RefMap()
    assert rm.available(), "build oracle/_ref first (make -C oracle ref)"
    g = {}
    # reprojector
    seed, max_fts = 5, 120
    sc = ms.build_map_scene(oracle, seed=seed)
    cfg = sc["cfg"]
    rm.config(cfg["n_pyr"], sc["cell"], max_fts)
    r = rm.reproject_map(sc["kf_imgs"], sc["T_kf"], sc["cur_img"], sc["T_cur"], cam_of(cfg), sc["points"], sc["obs"], sc["n_candidates"])
    g.update(reproj_seed=seed, reproj_max_fts=max_fts, reproj_n_matches=r["n_matches"], reproj_n_trials=r["n_trials"],
             reproj_n_failed=r["n_failed"], reproj_n_succeeded=r["n_succeeded"], reproj_new_point=r["new_point"].astype(np.int64),
             reproj_new_px=r["new_px"], reproj_new_level=r["new_level"], reproj_new_type=r["new_type"], reproj_new_grad=r["new_grad"])
    # pose optimizer
    s = ms.pose_opt_scene(seed=3)
    img = np.zeros((s["cfg"]["h"], s["cfg"]["w"]), np.uint8)
    r = rm.pose_optimize(cam_of(s["cfg"]), img, s["px"], s["level"], s["pos"], s["T_init"])
    g.update(pose_seed=3, pose_T=r["T"], pose_A=r["A"], pose_outlier=r["outlier"], pose_num_obs=r["num_obs"],
             pose_scale_init_final=np.array([r["estimated_scale"], r["error_init"], r["error_final"]]))
    # point optimizer
    g["point_pos"] = np.array([rm.point_optimize(cam_of(ms.SMALL), img, p["T"], p["f"], p["pos0"]) for p in ms.point_opt_scene()])
    # input stage: the app's own YUV2RGB restated in the oracle, cv::cvtColor from cv2
    fr = ms.yuv_frame(64, 48, 1, 2, 0)
    rgba = om.yuv420_to_rgba(fr["y"], fr["u"], fr["v"], fr["uv_stride"], fr["uv_pixel_stride"], fr["w"], fr["h"], fr["y_stride"])
    g["yuv_gray"] = cv2.cvtColor(rgba, cv2.COLOR_RGBA2GRAY)
    g["cv2_version"] = np.array(cv2.__version__)
    # seed initialisation
    cfg = ms.SMALL
    img = synth.render(synth.make_texture(512), cfg, synth.trajectory(4, seed=2)[2])
    rng = np.random.RandomState(0)
    existing = np.c_[rng.uniform(0, cfg["w"], 40), rng.uniform(0, cfg["h"], 40)]
    rm.config(cfg["n_pyr"], 30, 120)
    xs, ys, lv, seeds = rm.initialize_seeds(cam_of(cfg), img, cfg["n_pyr"], 20, 8.0, existing, 2.2, 1.7)
    g["seedinit_xyl"] = np.stack([xs, ys, lv], 1)
    g["seedinit_seeds"] = seeds
    path = os.path.join(OUT, "map_golden.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
