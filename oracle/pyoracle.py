"""ctypes bindings for the CPU oracle (oracle/svo_oracle.c) and, when it was built in a
container that has /root/reference, the real reference (oracle/_ref/libsvo_ref*.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under android_svo_b200/ imports this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
MAX_LEVELS = 8

c_u8p = C.POINTER(C.c_uint8)
c_dp = C.POINTER(C.c_double)
c_fp = C.POINTER(C.c_float)
c_ip = C.POINTER(C.c_int)


def _p(a, t):
    return a.ctypes.data_as(t)


def u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def build(force=False):
    """make oracle (+ ref when the reference tree exists). Safe to call repeatedly."""
    need = force or not os.path.exists(os.path.join(OUT, "libsvo_oracle.so"))
    src_newer = False
    if not need:
        so_t = os.path.getmtime(os.path.join(OUT, "libsvo_oracle.so"))
        for f in ("svo_oracle.c", "svo_oracle.h", "svo_cpu_pipeline.c", "svo_oracle_map.c"):
            if os.path.getmtime(os.path.join(HERE, f)) > so_t:
                src_newer = True
    if need or src_newer:
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if os.path.isdir("/root/reference/app/src/main/cpp/svo"):
        ref_so = os.path.join(OUT, "libsvo_ref.so")
        stale = not os.path.exists(ref_so) or any(
            os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(ref_so)
            for f in ("ref_harness.cpp", "ref_harness_map.cpp", "shim/cv_fast.cpp", "shim/opencv2/opencv.hpp", "svo_oracle.c"))
        stale = stale or not os.path.exists(os.path.join(OUT, "libsvo_ref_o3.so"))
        if force or stale:
            subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
        # the same harness over the C++ drop-in (needs the product library to link against)
        if os.path.exists(os.path.join(HERE, "..", "android_svo_b200", "lib", "libsvob200.so")):
            subprocess.check_call(["make", "-s", "-C", HERE, "dropin"])


class Cam(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("fx", C.c_double), ("fy", C.c_double),
                ("cx", C.c_double), ("cy", C.c_double)]

    @staticmethod
    def make(w, h, fx, fy, cx, cy):
        return Cam(int(w), int(h), float(fx), float(fy), float(cx), float(cy))

    def wh(self):
        return i32([self.width, self.height])

    def k(self):
        return f64([self.fx, self.fy, self.cx, self.cy])


class Pyr(C.Structure):
    _fields_ = [("data", c_u8p * MAX_LEVELS), ("w", C.c_int * MAX_LEVELS), ("h", C.c_int * MAX_LEVELS),
                ("n_levels", C.c_int)]


class Corner(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("level", C.c_int), ("score", C.c_float)]


class AlignOpts(C.Structure):
    _fields_ = [("max_level", C.c_int), ("min_level", C.c_int), ("n_iter", C.c_int), ("eps", C.c_double)]


class AlignResult(C.Structure):
    _fields_ = [("T_cur_ref", C.c_double * 7), ("H", C.c_double * 36), ("Jres", C.c_double * 6), ("x", C.c_double * 6),
                ("chi2", C.c_double), ("n_meas", C.c_int), ("iters", C.c_int * MAX_LEVELS), ("stop", C.c_int),
                ("n_ambiguous", C.c_int)]


class MatcherOpts(C.Structure):
    _fields_ = [("align_1d", C.c_int), ("align_max_iter", C.c_int), ("max_epi_search_steps", C.c_int),
                ("subpix_refinement", C.c_int), ("epi_search_edgelet_filtering", C.c_int),
                ("epi_search_edgelet_max_angle", C.c_double), ("max_search_level", C.c_int)]


class RefFeature(C.Structure):
    _fields_ = [("px_ref", C.c_double * 2), ("f_ref", C.c_double * 3), ("level_ref", C.c_int), ("type", C.c_int),
                ("grad", C.c_double * 2)]


class MatchResult(C.Structure):
    _fields_ = [("success", C.c_int), ("search_level", C.c_int), ("A_cur_ref", C.c_double * 4), ("h_inv", C.c_double),
                ("px_cur", C.c_double * 2), ("patch_with_border", C.c_uint8 * 100), ("patch", C.c_uint8 * 64)]


class EpiResult(C.Structure):
    _fields_ = [("success", C.c_int), ("depth", C.c_double), ("px_cur", C.c_double * 2), ("epi_length", C.c_double),
                ("search_level", C.c_int), ("reject", C.c_int), ("zmssd_best", C.c_int), ("n_evals", C.c_int),
                ("n_steps", C.c_int), ("A_cur_ref", C.c_double * 4), ("h_inv", C.c_double),
                ("patch_with_border", C.c_uint8 * 100), ("patch", C.c_uint8 * 64)]


class Seed(C.Structure):
    _fields_ = [("a", C.c_float), ("b", C.c_float), ("mu", C.c_float), ("z_range", C.c_float), ("sigma2", C.c_float)]


SEED_BEHIND, SEED_NOT_IN_FRAME, SEED_NO_MATCH, SEED_UPDATED, SEED_CONVERGED, SEED_NAN_ERASED = 1, 2, 3, 4, 5, 6


class Pyramid:
    """Dense host pyramid (stride == w) + the svo_pyr view over it."""

    def __init__(self, levels):
        self.levels = [u8(l) for l in levels]
        self.c = Pyr()
        self.c.n_levels = len(self.levels)
        for i, l in enumerate(self.levels):
            self.c.data[i] = _p(l, c_u8p)
            self.c.w[i] = l.shape[1]
            self.c.h[i] = l.shape[0]

    def __len__(self):
        return len(self.levels)

    def __getitem__(self, i):
        return self.levels[i]


class Oracle:
    """The C restatement."""

    def __init__(self):
        build()
        self.lib = L = C.CDLL(os.path.join(OUT, "libsvo_oracle.so"))
        L.svo_oracle_pyramid_bytes.restype = C.c_size_t
        L.svo_oracle_shi_tomasi.restype = C.c_float
        L.svo_oracle_compute_tau.restype = C.c_double
        L.svo_oracle_compute_tau.argtypes = [c_dp, c_dp, C.c_double, C.c_double]
        L.svo_oracle_update_seed.argtypes = [C.c_float, C.c_float, C.POINTER(Seed)]
        L.svo_oracle_seed_init.argtypes = [C.POINTER(Seed), C.c_float, C.c_float]
        L.svo_oracle_cam2world.argtypes = [C.POINTER(Cam), C.c_double, C.c_double, c_dp]
        L.svo_oracle_warp_matrix_affine.argtypes = [C.POINTER(Cam), C.POINTER(Cam), c_dp, c_dp, C.c_double, c_dp, C.c_int, c_dp]
        L.svo_oracle_find_match_direct.argtypes = [C.POINTER(Pyr), C.POINTER(Pyr), C.POINTER(Cam), C.POINTER(RefFeature),
                                                   C.c_double, c_dp, C.POINTER(MatcherOpts), c_dp, C.POINTER(MatchResult)]
        L.svo_oracle_find_epipolar_match.argtypes = [C.POINTER(Pyr), C.POINTER(Pyr), C.POINTER(Cam), C.POINTER(RefFeature),
                                                     c_dp, C.c_double, C.c_double, C.c_double, C.POINTER(MatcherOpts),
                                                     C.POINTER(EpiResult)]
        L.svo_oracle_update_seed_with_frame.argtypes = [C.POINTER(Pyr), C.POINTER(Pyr), C.POINTER(Cam), C.POINTER(RefFeature),
                                                        c_dp, c_dp, C.POINTER(MatcherOpts), C.c_double, C.POINTER(Seed),
                                                        C.POINTER(EpiResult)]
        L.svo_oracle_fast_detect.argtypes = [C.POINTER(Pyr), C.c_int, C.c_int, C.c_double, c_u8p, C.POINTER(Corner)]

    # -- pyramid
    def half_sample(self, img, mode):
        img = u8(img)
        h, w = img.shape
        out = np.zeros((h // 2, w // 2), np.uint8)
        self.lib.svo_oracle_half_sample(_p(img, c_u8p), w, h, _p(out, c_u8p), int(mode))
        return out

    def pyramid(self, img, n_levels, modes=None):
        img = u8(img)
        h, w = img.shape
        buf = np.zeros(self.lib.svo_oracle_pyramid_bytes(w, h, n_levels), np.uint8)
        m = _p(i32(modes), c_ip) if modes is not None else None
        self.lib.svo_oracle_build_pyramid(_p(img, c_u8p), w, h, n_levels, m, _p(buf, c_u8p))
        levels, off = [img], 0
        for _ in range(1, n_levels):
            w, h = w // 2, h // 2
            levels.append(buf[off:off + w * h].reshape(h, w).copy())
            off += w * h
        return Pyramid(levels)

    def synth_render(self, tex, ppm, plane_z, cam, T_f_w):
        tex = u8(tex)
        out = np.zeros((cam.height, cam.width), np.uint8)
        self.lib.svo_oracle_synth_render.argtypes = [c_u8p, C.c_int, C.c_double, C.c_double, C.POINTER(Cam), c_dp, c_u8p]
        self.lib.svo_oracle_synth_render(_p(tex, c_u8p), tex.shape[0], float(ppm), float(plane_z), C.byref(cam), _p(f64(T_f_w), c_dp), _p(out, c_u8p))
        return out

    # -- FAST
    def fast(self, img, thr=10, nonmax=True):
        img = u8(img)
        h, w = img.shape
        cap = w * h
        xs, ys, sc = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        n = self.lib.svo_oracle_fast(_p(img, c_u8p), w, h, int(thr), int(nonmax), cap, _p(xs, c_ip), _p(ys, c_ip), _p(sc, c_ip))
        return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()

    def shi_tomasi(self, img, u, v):
        img = u8(img)
        return float(self.lib.svo_oracle_shi_tomasi(_p(img, c_u8p), img.shape[1], img.shape[0], int(u), int(v)))

    def fast_detect(self, pyr, n_detect_levels, cell, thr, occupancy=None):
        W, H = pyr[0].shape[1], pyr[0].shape[0]
        n_cells = -(-W // cell) * -(-H // cell)
        cells = (Corner * n_cells)()
        occ = _p(u8(occupancy), c_u8p) if occupancy is not None else None
        n = self.lib.svo_oracle_fast_detect(C.byref(pyr.c), n_detect_levels, cell, float(thr), occ, cells)
        arr = np.array([(c.x, c.y, c.level, c.score) for c in cells],
                       dtype=[("x", "i4"), ("y", "i4"), ("level", "i4"), ("score", "f4")])
        return n, arr

    # -- SE3 / camera
    def se3_mul(self, A, B):
        o = np.zeros(7)
        self.lib.svo_oracle_se3_mul(_p(f64(A), c_dp), _p(f64(B), c_dp), _p(o, c_dp))
        return o

    def se3_inverse(self, A):
        o = np.zeros(7)
        self.lib.svo_oracle_se3_inverse(_p(f64(A), c_dp), _p(o, c_dp))
        return o

    def se3_exp(self, x):
        o = np.zeros(7)
        self.lib.svo_oracle_se3_exp(_p(f64(x), c_dp), _p(o, c_dp))
        return o

    def se3_transform(self, T, p):
        o = np.zeros(3)
        self.lib.svo_oracle_se3_transform(_p(f64(T), c_dp), _p(f64(p), c_dp), _p(o, c_dp))
        return o

    def cam2world(self, cam, u, v):
        o = np.zeros(3)
        self.lib.svo_oracle_cam2world(C.byref(cam), float(u), float(v), _p(o, c_dp))
        return o

    # -- sparse align
    def sparse_align(self, ref, cur, cam, px, xyz_ref, has_point, T_init, max_level, min_level, n_iter=30, eps=1e-6):
        px, xyz_ref, has_point = f64(px), f64(xyz_ref), u8(has_point)
        N = len(has_point)
        opts = AlignOpts(max_level, min_level, n_iter, eps)
        res = AlignResult()
        n = self.lib.svo_oracle_sparse_align(C.byref(ref.c), C.byref(cur.c), C.byref(cam), N, _p(px, c_dp), _p(xyz_ref, c_dp),
                                             _p(has_point, c_u8p), _p(f64(T_init), c_dp), C.byref(opts), C.byref(res))
        return n, res

    # -- feature alignment
    def align2d(self, img, pwb, patch, n_iter, px):
        img = u8(img)
        p = f64(px).copy()
        ok = self.lib.svo_oracle_align2d(_p(img, c_u8p), img.shape[1], img.shape[0], _p(u8(pwb), c_u8p), _p(u8(patch), c_u8p),
                                         int(n_iter), _p(p, c_dp))
        return ok, p

    def align1d(self, img, dirv, pwb, patch, n_iter, px):
        img = u8(img)
        p = f64(px).copy()
        d = np.ascontiguousarray(dirv, dtype=np.float32)
        h_inv = C.c_double(0)
        ok = self.lib.svo_oracle_align1d(_p(img, c_u8p), img.shape[1], img.shape[0], _p(d, c_fp), _p(u8(pwb), c_u8p),
                                         _p(u8(patch), c_u8p), int(n_iter), _p(p, c_dp), C.byref(h_inv))
        return ok, p, h_inv.value

    # -- warp / zmssd
    def warp(self, ref_pyr, cam, px_ref, f_ref, depth_ref, T_cur_ref, level_ref, max_search_level):
        A = np.zeros(4)
        self.lib.svo_oracle_warp_matrix_affine(C.byref(cam), C.byref(cam), _p(f64(px_ref), c_dp), _p(f64(f_ref), c_dp),
                                               float(depth_ref), _p(f64(T_cur_ref), c_dp), int(level_ref), _p(A, c_dp))
        sl = self.lib.svo_oracle_best_search_level(_p(A, c_dp), int(max_search_level))
        patch = np.zeros(100, np.uint8)
        img = ref_pyr[level_ref]
        self.lib.svo_oracle_warp_affine(_p(A, c_dp), _p(img, c_u8p), img.shape[1], img.shape[0], _p(f64(px_ref), c_dp),
                                        int(level_ref), sl, 5, _p(patch, c_u8p))
        return A, sl, patch

    def zmssd(self, ref_patch, img, x, y):
        img = u8(img)
        rp = u8(ref_patch)
        off = (y - 4) * img.shape[1] + (x - 4)
        ptr = C.cast(C.c_void_p(int(img.ctypes.data) + int(off)), c_u8p)
        return self.lib.svo_oracle_zmssd(_p(rp, c_u8p), ptr, img.shape[1])

    def matcher_opts(self, n_pyr_levels, **kw):
        o = MatcherOpts()
        self.lib.svo_oracle_matcher_opts_default(C.byref(o), int(n_pyr_levels))
        for k, v in kw.items():
            setattr(o, k, v)
        return o

    @staticmethod
    def ref_feature(px, f, level, ftype=0, grad=(1.0, 0.0)):
        r = RefFeature()
        r.px_ref[0], r.px_ref[1] = float(px[0]), float(px[1])
        for i in range(3):
            r.f_ref[i] = float(f[i])
        r.level_ref, r.type = int(level), int(ftype)
        r.grad[0], r.grad[1] = float(grad[0]), float(grad[1])
        return r

    def find_match_direct(self, ref, cur, cam, ftr, depth_ref, T_cur_ref, opts, px_cur):
        r = MatchResult()
        ok = self.lib.svo_oracle_find_match_direct(C.byref(ref.c), C.byref(cur.c), C.byref(cam), C.byref(ftr), float(depth_ref),
                                                   _p(f64(T_cur_ref), c_dp), C.byref(opts), _p(f64(px_cur), c_dp), C.byref(r))
        return ok, r

    def find_epipolar_match(self, ref, cur, cam, ftr, T_cur_ref, d_est, d_min, d_max, opts):
        r = EpiResult()
        ok = self.lib.svo_oracle_find_epipolar_match(C.byref(ref.c), C.byref(cur.c), C.byref(cam), C.byref(ftr),
                                                     _p(f64(T_cur_ref), c_dp), float(d_est), float(d_min), float(d_max),
                                                     C.byref(opts), C.byref(r))
        return ok, r

    # -- depth filter
    def seed_init(self, depth_mean, depth_min):
        s = Seed()
        self.lib.svo_oracle_seed_init(C.byref(s), depth_mean, depth_min)
        return s

    def update_seed(self, x, tau2, seed):
        self.lib.svo_oracle_update_seed(x, tau2, C.byref(seed))

    def compute_tau(self, T_ref_cur, f, z, ang):
        return self.lib.svo_oracle_compute_tau(_p(f64(T_ref_cur), c_dp), _p(f64(f), c_dp), z, ang)

    def update_seed_with_frame(self, ref, cur, cam, ftr, T_ref_w, T_cur_w, opts, conv_thresh, seed, want_epi=False):
        epi = EpiResult()
        st = self.lib.svo_oracle_update_seed_with_frame(C.byref(ref.c), C.byref(cur.c), C.byref(cam), C.byref(ftr),
                                                        _p(f64(T_ref_w), c_dp), _p(f64(T_cur_w), c_dp), C.byref(opts),
                                                        float(conv_thresh), C.byref(seed), C.byref(epi))
        return (st, epi) if want_epi else st


class Ref:
    """The real reference, compiled for this host (oracle/_ref/libsvo_ref.so). available() is False on
    machines where it was never built (e.g. a checkout without /root/reference and without the
    prebuilt .so)."""

    def __init__(self, nosse=False, dropin=False, o3=False):
        # dropin: the SAME C++ harness linked over android_svo_b200/host/svo_b200_dropin.cpp instead of the
        # reference's hot-path TUs (oracle/_ref/libsvo_dropin.so) — every operator call lands in CUDA
        # o3: the timing-only -O3 / AVX2 / FMA build (never used for parity)
        name = "libsvo_dropin.so" if dropin else ("libsvo_ref_nosse.so" if nosse else ("libsvo_ref_o3.so" if o3 else "libsvo_ref.so"))
        self.path = os.path.join(OUT, name)
        self.lib = None
        if os.path.exists(self.path):
            self.lib = L = C.CDLL(self.path)
            L.svo_ref_shi_tomasi.restype = C.c_float
            L.svo_ref_compute_tau.restype = C.c_double
            L.svo_ref_compute_tau.argtypes = [c_dp, c_dp, C.c_double, C.c_double]
            L.svo_ref_update_seed.argtypes = [C.c_float, C.c_float, c_fp]
            L.svo_ref_seed_init.argtypes = [C.c_float, C.c_float, c_fp]
            L.svo_ref_cam2world.argtypes = [c_ip, c_dp, C.c_double, C.c_double, c_dp]
            L.svo_ref_fast_detect.argtypes = [c_u8p, c_ip, c_dp, C.c_int, C.c_int, C.c_double, C.c_int, c_dp, C.c_int, c_dp, c_ip]
            L.svo_ref_warp.argtypes = [c_u8p, C.c_int, C.c_int, c_ip, c_dp, c_dp, c_dp, C.c_double, c_dp, C.c_int, C.c_int,
                                       c_dp, c_ip, c_u8p]
            L.svo_ref_find_epipolar_match.argtypes = [c_u8p, c_u8p, c_ip, c_dp, c_dp, c_dp, c_dp, C.c_int, C.c_int, c_dp,
                                                      C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                                      c_dp, c_dp, c_dp, c_ip, c_ip, c_dp, c_dp, c_u8p, c_u8p]
            L.svo_ref_update_seeds.argtypes = [C.c_int, C.POINTER(c_u8p), c_dp, c_u8p, c_dp, c_ip, c_dp, C.c_int, c_ip, c_dp,
                                               c_ip, c_fp, c_ip, C.c_double]

    def available(self):
        return self.lib is not None

    def is_dropin(self):
        return bool(self.lib.svo_ref_is_dropin())

    def dropin_launches(self):
        self.lib.svo_ref_dropin_launches.restype = C.c_longlong
        return int(self.lib.svo_ref_dropin_launches())

    def config(self, n_pyr_levels, klt_max_level, klt_min_level=2):
        self.lib.svo_ref_config(n_pyr_levels, klt_max_level, klt_min_level)

    def half_sample(self, img):
        img = u8(img)
        h, w = img.shape
        out = np.zeros((h // 2, w // 2), np.uint8)
        self.lib.svo_ref_half_sample(_p(img, c_u8p), w, h, _p(out, c_u8p))
        return out

    def pyramid(self, img, n_levels):
        img = u8(img)
        h, w = img.shape
        n = sum((w >> l) * (h >> l) for l in range(1, n_levels))
        buf = np.zeros(n, np.uint8)
        self.lib.svo_ref_pyramid(_p(img, c_u8p), w, h, n_levels, _p(buf, c_u8p))
        levels, off = [img], 0
        for _ in range(1, n_levels):
            w, h = w // 2, h // 2
            levels.append(buf[off:off + w * h].reshape(h, w).copy())
            off += w * h
        return Pyramid(levels)

    def shi_tomasi(self, img, u, v):
        img = u8(img)
        return float(self.lib.svo_ref_shi_tomasi(_p(img, c_u8p), img.shape[1], img.shape[0], int(u), int(v)))

    def zmssd(self, ref_patch, img, x, y):
        img = u8(img)
        off = (y - 4) * img.shape[1] + (x - 4)
        ptr = C.cast(C.c_void_p(int(img.ctypes.data) + int(off)), c_u8p)
        return self.lib.svo_ref_zmssd(_p(u8(ref_patch), c_u8p), ptr, img.shape[1])

    def fast_detect(self, img, cam, n_detect_levels, cell, thr, occ_px=None):
        img = u8(img)
        cap = 1 << 16
        px = np.zeros(2 * cap)
        lv = np.zeros(cap, np.int32)
        occ = f64(occ_px if occ_px is not None else np.zeros((0, 2)))
        n = self.lib.svo_ref_fast_detect(_p(img, c_u8p), _p(cam.wh(), c_ip), _p(cam.k(), c_dp), n_detect_levels, cell, float(thr),
                                         len(occ), _p(occ, c_dp), cap, _p(px, c_dp), _p(lv, c_ip))
        return px[:2 * n].reshape(n, 2).copy(), lv[:n].copy()

    def align2d(self, img, pwb, patch, n_iter, px):
        img = u8(img)
        p = f64(px).copy()
        ok = self.lib.svo_ref_align2d(_p(img, c_u8p), img.shape[1], img.shape[0], _p(u8(pwb), c_u8p), _p(u8(patch), c_u8p),
                                      int(n_iter), _p(p, c_dp))
        return ok, p

    def align1d(self, img, dirv, pwb, patch, n_iter, px):
        img = u8(img)
        p = f64(px).copy()
        d = np.ascontiguousarray(dirv, dtype=np.float32)
        h_inv = C.c_double(0)
        ok = self.lib.svo_ref_align1d(_p(img, c_u8p), img.shape[1], img.shape[0], _p(d, c_fp), _p(u8(pwb), c_u8p),
                                      _p(u8(patch), c_u8p), int(n_iter), _p(p, c_dp), C.byref(h_inv))
        return ok, p, h_inv.value

    def se3_mul(self, A, B):
        o = np.zeros(7)
        self.lib.svo_ref_se3_mul(_p(f64(A), c_dp), _p(f64(B), c_dp), _p(o, c_dp))
        return o

    def se3_inverse(self, A):
        o = np.zeros(7)
        self.lib.svo_ref_se3_inverse(_p(f64(A), c_dp), _p(o, c_dp))
        return o

    def se3_exp(self, x):
        o = np.zeros(7)
        self.lib.svo_ref_se3_exp(_p(f64(x), c_dp), _p(o, c_dp))
        return o

    def se3_transform(self, T, p):
        o = np.zeros(3)
        self.lib.svo_ref_se3_transform(_p(f64(T), c_dp), _p(f64(p), c_dp), _p(o, c_dp))
        return o

    def cam2world(self, cam, u, v):
        o = np.zeros(3)
        self.lib.svo_ref_cam2world(_p(cam.wh(), c_ip), _p(cam.k(), c_dp), float(u), float(v), _p(o, c_dp))
        return o

    def sparse_align(self, ref_img, cur_img, cam, max_level, min_level, n_iter, T_ref_w, T_cur_w_init, px, pt_world, level=None):
        ref_img, cur_img = u8(ref_img), u8(cur_img)
        px, pt_world = f64(px), f64(pt_world)
        N = len(px) // 2 if px.ndim == 1 else px.shape[0]
        lv = _p(i32(level), c_ip) if level is not None else None
        out = dict(T_cur_w=np.zeros(7), T_cur_ref_init=np.zeros(7), T_cur_ref=np.zeros(7), H=np.zeros(36), Jres=np.zeros(6),
                   x=np.zeros(6), f=np.zeros(3 * N), xyz_ref=np.zeros(3 * N), iters=np.zeros(8, np.int32))
        chi2, n_meas, stop = C.c_double(0), C.c_int(0), C.c_int(0)
        n = self.lib.svo_ref_sparse_align(_p(ref_img, c_u8p), _p(cur_img, c_u8p), _p(cam.wh(), c_ip), _p(cam.k(), c_dp),
                                          int(max_level), int(min_level), int(n_iter), _p(f64(T_ref_w), c_dp),
                                          _p(f64(T_cur_w_init), c_dp), N, _p(px, c_dp), lv, _p(pt_world, c_dp),
                                          _p(out["T_cur_w"], c_dp), _p(out["T_cur_ref_init"], c_dp), _p(out["T_cur_ref"], c_dp),
                                          _p(out["H"], c_dp), _p(out["Jres"], c_dp), _p(out["x"], c_dp), C.byref(chi2),
                                          C.byref(n_meas), _p(out["iters"], c_ip), C.byref(stop), _p(out["f"], c_dp),
                                          _p(out["xyz_ref"], c_dp))
        out.update(chi2=chi2.value, n_meas=n_meas.value, stop=stop.value, n_tracked=n)
        return out

    def warp(self, ref_level_img, cam, px_ref, f_ref, depth_ref, T_cur_ref, level_ref, max_search_level):
        img = u8(ref_level_img)
        A = np.zeros(4)
        sl = C.c_int(0)
        patch = np.zeros(100, np.uint8)
        self.lib.svo_ref_warp(_p(img, c_u8p), img.shape[1], img.shape[0], _p(cam.wh(), c_ip), _p(cam.k(), c_dp),
                              _p(f64(px_ref), c_dp), _p(f64(f_ref), c_dp), float(depth_ref), _p(f64(T_cur_ref), c_dp),
                              int(level_ref), int(max_search_level), _p(A, c_dp), C.byref(sl), _p(patch, c_u8p))
        return A, sl.value, patch

    def find_match_direct(self, imgs, T_f_w, cam, pt_world, obs_frame, obs_px, obs_level, cur_frame, px_cur,
                          obs_type=None, obs_grad=None):
        imgs = [u8(i) for i in imgs]
        arr = (c_u8p * len(imgs))(*[_p(i, c_u8p) for i in imgs])
        n_obs = len(obs_frame)
        obs_type = i32(obs_type if obs_type is not None else np.zeros(n_obs))
        obs_grad = f64(obs_grad if obs_grad is not None else np.tile([1.0, 0.0], n_obs))
        p = f64(px_cur).copy()
        chosen, sl, h_inv = C.c_int(-1), C.c_int(-1), C.c_double(0)
        A = np.zeros(4)
        pwb, patch = np.zeros(100, np.uint8), np.zeros(64, np.uint8)
        ok = self.lib.svo_ref_find_match_direct(len(imgs), arr, _p(f64(T_f_w), c_dp), _p(cam.wh(), c_ip), _p(cam.k(), c_dp),
                                                _p(f64(pt_world), c_dp), n_obs, _p(i32(obs_frame), c_ip), _p(f64(obs_px), c_dp),
                                                _p(i32(obs_level), c_ip), _p(obs_type, c_ip), _p(obs_grad, c_dp), int(cur_frame),
                                                _p(p, c_dp), C.byref(chosen), C.byref(sl), _p(A, c_dp), C.byref(h_inv),
                                                _p(pwb, c_u8p), _p(patch, c_u8p))
        return dict(success=ok, px_cur=p, chosen=chosen.value, search_level=sl.value, A=A, h_inv=h_inv.value, pwb=pwb, patch=patch)

    def find_epipolar_match(self, ref_img, cur_img, cam, T_ref_w, T_cur_w, px_ref, level_ref, d_est, d_min, d_max,
                            ftype=0, grad=(1.0, 0.0), align_1d=0, align_max_iter=10, max_epi_search_steps=1000, subpix=1):
        ref_img, cur_img = u8(ref_img), u8(cur_img)
        depth, epi_len = C.c_double(0), C.c_double(0)
        sl, rej = C.c_int(0), C.c_int(0)
        px_cur, A, f_ref = np.zeros(2), np.zeros(4), np.zeros(3)
        pwb, patch = np.zeros(100, np.uint8), np.zeros(64, np.uint8)
        ok = self.lib.svo_ref_find_epipolar_match(_p(ref_img, c_u8p), _p(cur_img, c_u8p), _p(cam.wh(), c_ip), _p(cam.k(), c_dp),
                                                  _p(f64(T_ref_w), c_dp), _p(f64(T_cur_w), c_dp), _p(f64(px_ref), c_dp),
                                                  int(level_ref), int(ftype), _p(f64(grad), c_dp), float(d_est), float(d_min),
                                                  float(d_max), int(align_1d), int(align_max_iter), int(max_epi_search_steps),
                                                  int(subpix), C.byref(depth), _p(px_cur, c_dp), C.byref(epi_len), C.byref(sl),
                                                  C.byref(rej), _p(A, c_dp), _p(f_ref, c_dp), _p(pwb, c_u8p), _p(patch, c_u8p))
        return dict(success=ok, depth=depth.value, px_cur=px_cur, epi_length=epi_len.value, search_level=sl.value,
                    reject=rej.value, A=A, f_ref=f_ref, pwb=pwb, patch=patch)

    def seed_init(self, depth_mean, depth_min):
        s = np.zeros(5, np.float32)
        self.lib.svo_ref_seed_init(depth_mean, depth_min, _p(s, c_fp))
        return s

    def update_seed(self, x, tau2, state):
        s = np.ascontiguousarray(state, dtype=np.float32).copy()
        self.lib.svo_ref_update_seed(x, tau2, _p(s, c_fp))
        return s

    def compute_tau(self, T_ref_cur, f, z, ang):
        return self.lib.svo_ref_compute_tau(_p(f64(T_ref_cur), c_dp), _p(f64(f), c_dp), z, ang)

    def update_seeds(self, ref_imgs, T_ref_w, cur_img, T_cur_w, cam, seed_ref, px, level, state, conv_thresh=100.0):
        ref_imgs = [u8(i) for i in ref_imgs]
        arr = (c_u8p * len(ref_imgs))(*[_p(i, c_u8p) for i in ref_imgs])
        cur_img = u8(cur_img)
        S = len(seed_ref)
        st = np.ascontiguousarray(state, dtype=np.float32).copy()
        status = np.zeros(S, np.int32)
        self.lib.svo_ref_update_seeds(len(ref_imgs), arr, _p(f64(T_ref_w), c_dp), _p(cur_img, c_u8p), _p(f64(T_cur_w), c_dp),
                                      _p(cam.wh(), c_ip), _p(cam.k(), c_dp), S, _p(i32(seed_ref), c_ip), _p(f64(px), c_dp),
                                      _p(i32(level), c_ip), _p(st, c_fp), _p(status, c_ip), float(conv_thresh))
        return st.reshape(S, 5), status


class StepStats(C.Structure):
    _fields_ = [("T_cur_w", C.c_double * 7), ("chi2", C.c_double), ("n_tracked", C.c_int), ("n_matched", C.c_int),
                ("n_seeds_updated", C.c_int), ("n_seeds_converged", C.c_int), ("n_seeds_failed", C.c_int),
                ("n_seeds_skipped", C.c_int), ("align_iters", C.c_int), ("n_exact_chi2", C.c_int),
                ("n_reproj_trials", C.c_int), ("n_pose_obs", C.c_int)]


class _SeqBase:
    """One sequence stepped frame by frame: pyramid -> sparse align -> reprojection refinement -> seed update."""

    def step(self, cur_img, T_last_w, last_px, want_px=False):
        cur_img = u8(cur_img)
        last_px = f64(last_px)
        st = StepStats()
        n = self.N
        px = np.zeros((n, 2)) if want_px else None
        ok = np.zeros(n, np.int32) if want_px else None
        self._step(self.h, _p(cur_img, c_u8p), _p(f64(T_last_w), c_dp), _p(last_px, c_dp), C.byref(st),
                   _p(px, c_dp) if want_px else None, _p(ok, c_ip) if want_px else None)
        return (st, px, ok) if want_px else st


class OracleSeq(_SeqBase):
    def __init__(self, oracle, cam, n_levels, max_level, min_level, n_pyr_levels_cfg, conv_thresh=100.0, depth_mean=2.4,
                 depth_min=1.2, reseed=1, n_iter=30):
        self.lib = L = oracle.lib
        L.svo_oracle_seq_create.restype = C.c_void_p
        L.svo_oracle_seq_create.argtypes = [C.POINTER(Cam), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_float, C.c_float, C.c_int]
        L.svo_oracle_seq_destroy.argtypes = [C.c_void_p]
        L.svo_oracle_seq_set_keyframe.argtypes = [C.c_void_p, c_u8p, c_dp, C.c_int, c_dp, c_ip, c_dp, C.c_int, c_dp, c_ip]
        L.svo_oracle_seq_set_last.argtypes = [C.c_void_p, c_u8p]
        L.svo_oracle_seq_step.argtypes = [C.c_void_p, c_u8p, c_dp, c_dp, C.POINTER(StepStats), c_dp, c_ip]
        L.svo_oracle_seq_get_seeds.argtypes = [C.c_void_p, C.POINTER(Seed)]
        self.h = L.svo_oracle_seq_create(C.byref(cam), n_levels, max_level, min_level, n_iter, n_pyr_levels_cfg, conv_thresh,
                                         depth_mean, depth_min, reseed)
        self._step = L.svo_oracle_seq_step
        self.N = self.S = 0

    def set_keyframe(self, img, T_kf_w, kf_px, kf_level, pt_world, seed_px, seed_level):
        self.N, self.S = len(kf_level), len(seed_level)
        self.lib.svo_oracle_seq_set_keyframe(self.h, _p(u8(img), c_u8p), _p(f64(T_kf_w), c_dp), self.N, _p(f64(kf_px), c_dp),
                                             _p(i32(kf_level), c_ip), _p(f64(pt_world), c_dp), self.S, _p(f64(seed_px), c_dp),
                                             _p(i32(seed_level), c_ip))

    def set_last(self, img):
        self.lib.svo_oracle_seq_set_last(self.h, _p(u8(img), c_u8p))

    def set_chain(self, cell_size=30, max_fts=120, pose_opt=1):
        """Reprojector::reprojectMap + pose optimiser between alignment and the depth filter (frame_handler_mono.cpp:191-222)"""
        self.lib.svo_oracle_seq_set_chain.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        self.lib.svo_oracle_seq_set_chain(self.h, int(cell_size), int(max_fts), int(pose_opt))

    def num_slots(self):
        self.lib.svo_oracle_seq_num_slots.argtypes = [C.c_void_p]
        return int(self.lib.svo_oracle_seq_num_slots(self.h))

    def seeds(self):
        self.S = self.num_slots()
        arr = (Seed * max(self.S, 1))()
        self.lib.svo_oracle_seq_get_seeds(self.h, arr)
        return np.array([(s.a, s.b, s.mu, s.z_range, s.sigma2) for s in arr[:self.S]], np.float32).reshape(self.S, 5)

    # ---- keyframe insertion (DepthFilter::addKeyframe -> initializeSeeds, ageing, removeKeyframe)
    def set_pool(self, max_kfs=4, max_n_kfs=3, reseed=3):
        self.lib.svo_oracle_seq_set_pool.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        self.lib.svo_oracle_seq_set_pool(self.h, int(max_kfs), int(max_n_kfs), int(reseed))

    def set_detector(self, cell, levels, thr):
        self.lib.svo_oracle_seq_set_detector.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        self.lib.svo_oracle_seq_set_detector(self.h, int(cell), int(levels), float(thr))

    def add_keyframe(self, depth_mean, depth_min):
        self.lib.svo_oracle_seq_add_keyframe.argtypes = [C.c_void_p, C.c_float, C.c_float]
        n = int(self.lib.svo_oracle_seq_add_keyframe(self.h, float(depth_mean), float(depth_min)))
        self.S = self.num_slots()
        return n

    def seed_refs(self):
        """(px[S,2], level, kf, batch_id, state) of every slot; state 0 = alive"""
        S = self.num_slots()
        px = np.zeros((max(S, 1), 2)); lv = np.zeros(max(S, 1), np.int32); kf = np.zeros(max(S, 1), np.int32)
        bt = np.zeros(max(S, 1), np.int32); st = np.zeros(max(S, 1), np.int32)
        self.lib.svo_oracle_seq_get_seed_refs.argtypes = [C.c_void_p, c_dp, c_ip, c_ip, c_ip, c_ip]
        self.lib.svo_oracle_seq_get_seed_refs(self.h, _p(px, c_dp), _p(lv, c_ip), _p(kf, c_ip), _p(bt, c_ip), _p(st, c_ip))
        return px[:S], lv[:S], kf[:S], bt[:S], st[:S]

    def seed_obs(self):
        """per-seed observation of the last step, same fields as svob200_seed_obs"""
        dt = np.dtype([("status", "i4"), ("search_level", "i4"), ("zmssd_best", "i4"), ("n_evals", "i4"), ("z", "f8"),
                       ("px_cur", "f8", 2), ("epi_length", "f8")], align=True)
        self.S = self.num_slots()
        out = np.zeros(max(self.S, 1), dt)
        self.lib.svo_oracle_seq_get_seed_obs.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.svo_oracle_seq_get_seed_obs(self.h, out.ctypes.data)
        return out[:self.S]

    def set_pose_override(self, T_cur_w):
        """the matcher / depth-filter stages of the NEXT step use this pose instead of the aligned one"""
        self.lib.svo_oracle_seq_set_pose_override.argtypes = [C.c_void_p, c_dp]
        self.lib.svo_oracle_seq_set_pose_override(self.h, _p(f64(T_cur_w), c_dp))

    def close(self):
        if self.h:
            self.lib.svo_oracle_seq_destroy(self.h)
            self.h = None


class RefSeq(_SeqBase):
    """Same step through the real reference classes (oracle/_ref/libsvo_ref.so)."""

    def __init__(self, ref, cam, n_levels, max_level, min_level, n_pyr_levels_cfg, conv_thresh=100.0, depth_mean=2.4,
                 depth_min=1.2, reseed=1, n_iter=30):
        self.lib = L = ref.lib
        L.svo_ref_seq_create.restype = C.c_void_p
        L.svo_ref_seq_create.argtypes = [c_ip, c_dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_float, C.c_float, C.c_int]
        L.svo_ref_seq_destroy.argtypes = [C.c_void_p]
        L.svo_ref_seq_set_keyframe.argtypes = [C.c_void_p, c_u8p, c_dp, C.c_int, c_dp, c_ip, c_dp, C.c_int, c_dp, c_ip]
        L.svo_ref_seq_set_last.argtypes = [C.c_void_p, c_u8p]
        L.svo_ref_seq_step.argtypes = [C.c_void_p, c_u8p, c_dp, c_dp, C.POINTER(StepStats), c_dp, c_ip]
        L.svo_ref_seq_get_seeds.argtypes = [C.c_void_p, c_fp]
        # Config is a process-wide singleton in the reference: pyramid depth = max(nPyrLevels, kltMaxLevel+1)
        ref.config(n_pyr_levels_cfg, n_levels - 1, min_level)
        self.h = L.svo_ref_seq_create(_p(cam.wh(), c_ip), _p(cam.k(), c_dp), max_level, min_level, n_iter, conv_thresh,
                                      depth_mean, depth_min, reseed)
        self._step = L.svo_ref_seq_step
        self.N = self.S = 0

    def set_keyframe(self, img, T_kf_w, kf_px, kf_level, pt_world, seed_px, seed_level):
        self.N, self.S = len(kf_level), len(seed_level)
        self.lib.svo_ref_seq_set_keyframe(self.h, _p(u8(img), c_u8p), _p(f64(T_kf_w), c_dp), self.N, _p(f64(kf_px), c_dp),
                                          _p(i32(kf_level), c_ip), _p(f64(pt_world), c_dp), self.S, _p(f64(seed_px), c_dp),
                                          _p(i32(seed_level), c_ip))

    def set_last(self, img):
        self.lib.svo_ref_seq_set_last(self.h, _p(u8(img), c_u8p))

    # ---- keyframe insertion through the reference's own DepthFilter::addKeyframe / removeKeyframe and FastDetector
    def set_pool(self, max_kfs=4, max_n_kfs=3, reseed=3, det_cell=30, det_levels=3, det_thr=20.0):
        """call before set_keyframe"""
        self.lib.svo_ref_seq_set_pool.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
        self.lib.svo_ref_seq_set_pool(self.h, int(max_kfs), int(max_n_kfs), int(reseed), int(det_cell), int(det_levels), float(det_thr))

    def add_keyframe(self, depth_mean, depth_min):
        self.lib.svo_ref_seq_add_keyframe.argtypes = [C.c_void_p, C.c_float, C.c_float]
        return int(self.lib.svo_ref_seq_add_keyframe(self.h, float(depth_mean), float(depth_min)))

    def seed_list(self, cap=100000):
        """the reference's std::list<Seed> as arrays: px[n,2], level, kf index, batch id, state[n,5]"""
        px = np.zeros((cap, 2)); lv = np.zeros(cap, np.int32); kf = np.zeros(cap, np.int32); bt = np.zeros(cap, np.int32)
        st = np.zeros((cap, 5), np.float32)
        self.lib.svo_ref_seq_get_seed_list.argtypes = [C.c_void_p, C.c_int, c_dp, c_ip, c_ip, c_ip, c_fp]
        n = int(self.lib.svo_ref_seq_get_seed_list(self.h, cap, _p(px, c_dp), _p(lv, c_ip), _p(kf, c_ip), _p(bt, c_ip), _p(st, c_fp)))
        return px[:n], lv[:n], kf[:n], bt[:n], st[:n]

    def set_threaded(self):
        """timing only: the depth filter in its own thread, like the app (call before set_keyframe)"""
        self.lib.svo_ref_seq_set_threaded.argtypes = [C.c_void_p]
        self.lib.svo_ref_seq_set_threaded(self.h)

    def drain(self):
        self.lib.svo_ref_seq_drain.argtypes = [C.c_void_p]
        self.lib.svo_ref_seq_drain(self.h)

    def timing(self):
        """steady_clock seconds per operator since the last call: dict(pyramid, align, refine, seeds, steps)"""
        out = np.zeros(5)
        self.lib.svo_ref_seq_get_timing.argtypes = [C.c_void_p, c_dp]
        self.lib.svo_ref_seq_get_timing(self.h, _p(out, c_dp))
        return dict(pyramid=out[0], align=out[1], refine=out[2], seeds=out[3], steps=int(out[4]))

    def set_chain(self, cell_size=30, max_fts=120, pose_opt=1):
        """the reference's own Reprojector + pose_optimizer between alignment and the depth filter"""
        self.lib.svo_ref_seq_set_chain.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        self.lib.svo_ref_seq_set_chain(self.h, int(cell_size), int(max_fts), int(pose_opt))

    def seeds(self):
        out = np.zeros((max(self.S, 1), 5), np.float32)
        self.lib.svo_ref_seq_get_seeds(self.h, _p(out, c_fp))
        return out[:self.S]

    def close(self):
        if self.h:
            self.lib.svo_ref_seq_destroy(self.h)
            self.h = None
