"""ctypes bindings for oracle/svo_oracle_map.c (C restatement) and oracle/ref_harness_map.cpp (the real
reference) — the callers either side of the hot path: reprojector, pose optimizer,
point optimizer, YUV->gray input stage, seed initialisation.

TEST INFRASTRUCTURE ONLY: imported by tests/ (and tests/golden/make_golden.py).  Nothing under
android_svo_b200/ imports this.
"""
import ctypes as C
import numpy as np

from .pyoracle import Oracle, Ref, Cam, Pyr, MatcherOpts, Corner, Seed, c_u8p, c_dp, c_fp, c_ip, _p, u8, f64, i32

POINT_DELETED, POINT_CANDIDATE, POINT_UNKNOWN, POINT_GOOD = 0, 1, 2, 3
REPROJ_NOT_IN_FRAME, REPROJ_UNTRIED, REPROJ_DELETED, REPROJ_FAILED, REPROJ_MATCHED = 0, 1, 2, 3, 4
TUKEY_B = 8.6851        # TukeyWeightFunction::DEFAULT_B (robust_cost.cpp:87)

map_point_dt = np.dtype([("pos", np.float64, 3), ("type", np.int32), ("obs_begin", np.int32), ("obs_end", np.int32)], align=True)
ref_feature_dt = np.dtype([("px_ref", np.float64, 2), ("f_ref", np.float64, 3), ("level_ref", np.int32), ("type", np.int32),
                           ("grad", np.float64, 2)], align=True)
point_obs_dt = np.dtype([("keyframe", np.int32), ("ftr", ref_feature_dt)], align=True)
reproj_result_dt = np.dtype([("status", np.int32), ("cell", np.int32), ("obs", np.int32), ("search_level", np.int32),
                             ("px", np.float64, 2), ("A_cur_ref", np.float64, 4)], align=True)
pose_opt_result_dt = np.dtype([("A", np.float64, 36), ("chi2", np.float64), ("estimated_scale", np.float64), ("error_init", np.float64),
                               ("error_final", np.float64), ("iters", np.int32), ("num_obs", np.int32), ("rolled_back", np.int32)],
                              align=True)
corner_dt = np.dtype([("x", np.int32), ("y", np.int32), ("level", np.int32), ("score", np.float32)], align=True)
seed_dt = np.dtype([("a", np.float32), ("b", np.float32), ("mu", np.float32), ("z_range", np.float32), ("sigma2", np.float32)], align=True)

vp = C.c_void_p


def _v(a):
    return a.ctypes.data_as(vp)


def n_cells(cam, cell):
    return int(np.ceil(cam.width / cell)) * int(np.ceil(cam.height / cell))


class OracleMap:
    """svo_oracle_map.c"""

    def __init__(self, oracle=None):
        self.o = oracle or Oracle()
        L = self.lib = self.o.lib
        L.svo_oracle_reproject_map.argtypes = [C.POINTER(C.POINTER(Pyr)), C.POINTER(Pyr), C.POINTER(Cam), c_dp, C.c_int, vp, vp, c_dp,
                                               C.c_int, C.c_int, C.POINTER(MatcherOpts), vp, c_ip, c_ip, c_ip]
        L.svo_oracle_pose_optimize.argtypes = [C.POINTER(Cam), C.c_int, c_dp, c_ip, c_dp, C.c_double, C.c_int, C.c_double, C.c_float,
                                               c_dp, vp, c_u8p]
        L.svo_oracle_point_optimize.argtypes = [C.c_int, c_dp, c_dp, C.c_int, C.c_double, c_dp]
        L.svo_oracle_yuv420_to_rgba.argtypes = [c_u8p, C.c_int, c_u8p, c_u8p, C.c_int, C.c_int, C.c_int, C.c_int, c_u8p]
        L.svo_oracle_yuv420_to_gray.argtypes = [c_u8p, C.c_int, c_u8p, c_u8p, C.c_int, C.c_int, C.c_int, C.c_int, c_u8p]
        L.svo_oracle_rgba_to_gray.argtypes = [c_u8p, C.c_int, C.c_int, c_u8p]
        L.svo_oracle_initialize_seeds.argtypes = [C.POINTER(Pyr), C.POINTER(Cam), C.c_int, C.c_int, C.c_double, C.c_int, c_dp,
                                                  C.c_float, C.c_float, vp, vp]
        L.svo_oracle_frame_pos.argtypes = [c_dp, c_dp]

    def frame_pos(self, T):
        out = np.zeros(3)
        self.lib.svo_oracle_frame_pos(_p(f64(T), c_dp), _p(out, c_dp))
        return out

    def reproject_map(self, kf_pyrs, cur_pyr, cam, T_cur_w, points, obs, T_kf_w, cell, max_fts, mopts):
        arr = (C.POINTER(Pyr) * len(kf_pyrs))(*[C.pointer(p.c) for p in kf_pyrs])
        points = np.ascontiguousarray(points, map_point_dt)
        obs = np.ascontiguousarray(obs, point_obs_dt)
        res = np.zeros(len(points), reproj_result_dt)
        winner = np.zeros(n_cells(cam, cell), np.int32)
        nm, nt = C.c_int(0), C.c_int(0)
        self.lib.svo_oracle_reproject_map(arr, C.byref(cur_pyr.c), C.byref(cam), _p(f64(T_cur_w), c_dp), len(points), _v(points), _v(obs),
                                          _p(f64(T_kf_w), c_dp), int(cell), int(max_fts), C.byref(mopts), _v(res), _p(winner, c_ip),
                                          C.byref(nm), C.byref(nt))
        return res, winner, nm.value, nt.value

    def pose_optimize(self, cam, f, level, pos, T_f_w, reproj_thresh=2.0, n_iter=10, eps=1e-10, tukey_b=TUKEY_B):
        f, pos, level = f64(f), f64(pos), i32(level)
        T = f64(T_f_w).copy()
        res = np.zeros(1, pose_opt_result_dt)
        outl = np.zeros(len(level), np.uint8)
        self.lib.svo_oracle_pose_optimize(C.byref(cam), len(level), _p(f, c_dp), _p(level, c_ip), _p(pos, c_dp), float(reproj_thresh),
                                          int(n_iter), float(eps), float(tukey_b), _p(T, c_dp), _v(res), _p(outl, c_u8p))
        return T, res[0], outl

    def point_optimize(self, T_f_w, f, pos, n_iter=20, eps=1e-10):
        T_f_w, f = f64(T_f_w), f64(f)
        p = f64(pos).copy()
        it = self.lib.svo_oracle_point_optimize(len(T_f_w), _p(T_f_w, c_dp), _p(f, c_dp), int(n_iter), float(eps), _p(p, c_dp))
        return p, it

    def yuv420_to_rgba(self, y, u, v, uv_stride, uv_pixel_stride, w, h, y_stride=None):
        y, u, v = u8(y), u8(u), u8(v)
        out = np.zeros((h, w, 4), np.uint8)
        self.lib.svo_oracle_yuv420_to_rgba(_p(y, c_u8p), int(y_stride or w), _p(u, c_u8p), _p(v, c_u8p), int(uv_stride), int(uv_pixel_stride),
                                           w, h, _p(out, c_u8p))
        return out

    def rgba_to_gray(self, rgba):
        rgba = u8(rgba)
        h, w = rgba.shape[:2]
        out = np.zeros((h, w), np.uint8)
        self.lib.svo_oracle_rgba_to_gray(_p(rgba, c_u8p), w, h, _p(out, c_u8p))
        return out

    def yuv420_to_gray(self, y, u, v, uv_stride, uv_pixel_stride, w, h, y_stride=None):
        y, u, v = u8(y), u8(u), u8(v)
        out = np.zeros((h, w), np.uint8)
        self.lib.svo_oracle_yuv420_to_gray(_p(y, c_u8p), int(y_stride or w), _p(u, c_u8p), _p(v, c_u8p), int(uv_stride), int(uv_pixel_stride),
                                           w, h, _p(out, c_u8p))
        return out

    def initialize_seeds(self, pyr, cam, n_detect_levels, cell, thr, existing_px, depth_mean, depth_min):
        epx = f64(existing_px).reshape(-1, 2)
        nc = n_cells(cam, cell)
        corners = np.zeros(nc, corner_dt)
        seeds = np.zeros(nc, seed_dt)
        n = self.lib.svo_oracle_initialize_seeds(C.byref(pyr.c), C.byref(cam), int(n_detect_levels), int(cell), float(thr), len(epx),
                                                 _p(epx, c_dp), float(depth_mean), float(depth_min), _v(corners), _v(seeds))
        return corners[:n], seeds[:n]


class RefMap:
    """ref_harness_map.cpp over the real reference (or, with dropin=True, over the B200 drop-in)."""

    def __init__(self, ref=None, **kw):
        self.r = ref or Ref(**kw)
        self.lib = self.r.lib
        if self.lib is not None and hasattr(self.lib, "svo_ref_reproject_map"):
            L = self.lib
            L.svo_ref_pose_optimize.argtypes = [c_ip, c_dp, c_u8p, C.c_int, c_dp, c_ip, c_dp, C.c_double, C.c_int, c_dp, c_dp, c_dp, c_ip, c_u8p]
            L.svo_ref_point_optimize.argtypes = [c_ip, c_dp, c_u8p, C.c_int, c_dp, c_dp, C.c_int, c_dp]
            L.svo_ref_initialize_seeds.argtypes = [c_ip, c_dp, c_u8p, C.c_int, C.c_int, C.c_double, C.c_int, c_dp, C.c_double, C.c_double,
                                                   C.c_int, c_ip, c_ip, c_ip, c_fp]
            L.svo_ref_optimize_structure.argtypes = [c_ip, c_dp, c_u8p, C.c_int, c_ip, c_ip, c_dp, c_dp, c_ip, C.c_int, C.c_int, c_dp, c_ip]

    def available(self):
        return self.lib is not None and hasattr(self.lib, "svo_ref_reproject_map")

    def config(self, n_pyr_levels, grid_size, max_fts, klt_max=4, klt_min=2):
        self.r.config(n_pyr_levels, klt_max, klt_min)
        self.lib.svo_ref_config_map(int(grid_size), int(max_fts))

    def reproject_map(self, kf_imgs, T_kf_w, cur_img, T_cur_w, cam, points, obs, n_candidates):
        """points: map_point_dt array; obs: point_obs_dt array (px/level/type/grad used; f is recomputed by the
        reference's Feature ctor).  Returns dict of the reference's observable outcome."""
        imgs = [u8(i) for i in kf_imgs]
        arr = (c_u8p * len(imgs))(*[_p(i, c_u8p) for i in imgs])
        n = len(points)
        pos = f64(points["pos"]); typ = i32(points["type"]); ob = i32(points["obs_begin"]); oe = i32(points["obs_end"])
        okf = i32(obs["keyframe"]); opx = f64(obs["ftr"]["px_ref"]); olv = i32(obs["ftr"]["level_ref"]); oty = i32(obs["ftr"]["type"])
        ogr = f64(obs["ftr"]["grad"])
        failed, succ, tafter = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
        cap = n + 8
        new_point, new_level, new_type = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        new_px, new_grad = np.zeros((cap, 2)), np.zeros((cap, 2))
        nm, nt = C.c_int(0), C.c_int(0)
        ov_kf, ov_cnt = np.zeros(len(imgs), np.int32), np.zeros(len(imgs), np.int32)
        cur = u8(cur_img)
        k = self.lib.svo_ref_reproject_map(len(imgs), arr, _p(f64(T_kf_w), c_dp), _p(cur, c_u8p), _p(f64(T_cur_w), c_dp), _p(cam.wh(), c_ip),
                                           _p(cam.k(), c_dp), n, int(n_candidates), _p(pos, c_dp), _p(typ, c_ip), _p(ob, c_ip), _p(oe, c_ip),
                                           _p(okf, c_ip), _p(opx, c_dp), _p(olv, c_ip), _p(oty, c_ip), _p(ogr, c_dp),
                                           _p(failed, c_ip), _p(succ, c_ip), _p(tafter, c_ip), cap, _p(new_point, c_ip), _p(new_px, c_dp),
                                           _p(new_level, c_ip), _p(new_type, c_ip), _p(new_grad, c_dp), C.byref(nm), C.byref(nt),
                                           _p(ov_kf, c_ip), _p(ov_cnt, c_ip))
        return dict(n_failed=failed, n_succeeded=succ, type_after=tafter, new_point=new_point[:k], new_px=new_px[:k], new_level=new_level[:k],
                    new_type=new_type[:k], new_grad=new_grad[:k], n_matches=nm.value, n_trials=nt.value, overlap_kf=ov_kf, overlap_cnt=ov_cnt)

    def pose_optimize(self, cam, img, px, level, pos, T_f_w, reproj_thresh=2.0, n_iter=10):
        px, pos, level = f64(px), f64(pos), i32(level)
        T = f64(T_f_w).copy()
        A = np.zeros(36); sif = np.zeros(3); nobs = C.c_int(0)
        outl = np.zeros(len(level), np.uint8)
        img = u8(img)
        self.lib.svo_ref_pose_optimize(_p(cam.wh(), c_ip), _p(cam.k(), c_dp), _p(img, c_u8p), len(level), _p(px, c_dp), _p(level, c_ip),
                                       _p(pos, c_dp), float(reproj_thresh), int(n_iter), _p(T, c_dp), _p(A, c_dp), _p(sif, c_dp),
                                       C.byref(nobs), _p(outl, c_u8p))
        return dict(T=T, A=A, estimated_scale=sif[0], error_init=sif[1], error_final=sif[2], num_obs=nobs.value, outlier=outl)

    def point_optimize(self, cam, img, T_f_w, f, pos, n_iter=20):
        T_f_w, f = f64(T_f_w), f64(f)
        p = f64(pos).copy()
        img = u8(img)
        self.lib.svo_ref_point_optimize(_p(cam.wh(), c_ip), _p(cam.k(), c_dp), _p(img, c_u8p), len(T_f_w), _p(T_f_w, c_dp), _p(f, c_dp),
                                        int(n_iter), _p(p, c_dp))
        return p

    def optimize_structure(self, cam, img, obs_offsets, T_f_w, f, pos, last_optim, max_n_pts, n_iter):
        off = i32(obs_offsets)
        n = len(off) - 1
        ob, oe = i32(off[:-1]), i32(off[1:])
        p = f64(pos).reshape(n, 3).copy()
        done = np.zeros(n, np.int32)
        img = u8(img)
        self.lib.svo_ref_optimize_structure(_p(cam.wh(), c_ip), _p(cam.k(), c_dp), _p(img, c_u8p), n, _p(ob, c_ip), _p(oe, c_ip),
                                            _p(f64(T_f_w), c_dp), _p(f64(f), c_dp), _p(i32(last_optim), c_ip), int(max_n_pts), int(n_iter),
                                            _p(p, c_dp), _p(done, c_ip))
        return p, done

    def initialize_seeds(self, cam, img, n_detect_levels, cell, thr, existing_px, depth_mean, depth_min):
        epx = f64(existing_px).reshape(-1, 2)
        cap = n_cells(cam, cell)
        xs, ys, lv = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        seeds = np.zeros((cap, 5), np.float32)
        img = u8(img)
        n = self.lib.svo_ref_initialize_seeds(_p(cam.wh(), c_ip), _p(cam.k(), c_dp), _p(img, c_u8p), int(n_detect_levels), int(cell), float(thr),
                                              len(epx), _p(epx, c_dp), float(depth_mean), float(depth_min), cap, _p(xs, c_ip), _p(ys, c_ip),
                                              _p(lv, c_ip), _p(seeds, c_fp))
        return xs[:n], ys[:n], lv[:n], seeds[:n]
