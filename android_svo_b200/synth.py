"""Deterministic synthetic sequences for the SVO front end.

A fronto-parallel textured plane at z = plane_z seen by a distortion-free pinhole camera that
moves on a small sinusoidal trajectory.  Everything here is integer or float64 numpy and fully
seeded, so the same bytes reach the CUDA path, the C oracle and the compiled reference.
Rendering is outside every timed region.

Pose layout: float64[7] = (tx,ty,tz,qx,qy,qz,qw), the reference's SE3 constructor order
(reference: svo/include/svo/SE3.h:17-19); T_f_w maps world points into the camera frame.
"""
import numpy as np

CONFIGS = {
    # name: (width, height, fx, fy, cx, cy, n_levels, n_features, n_seeds, align_max_level, align_min_level)
    # n_levels = pyramid depth = max(nPyrLevels, kltMaxLevel+1) (frame.cpp:63); n_pyr = Config::nPyrLevels()
    # (detector levels and search-level cap); max/min_level = kltMaxLevel / kltMinLevel
    "C2": dict(w=640, h=480, fx=525.0, fy=525.0, cx=319.5, cy=239.5, n_levels=4, n_pyr=4, n_features=120, n_seeds=768,
               max_level=3, min_level=2),
    "C3": dict(w=752, h=480, fx=458.0, fy=458.0, cx=367.2, cy=248.4, n_levels=5, n_pyr=5, n_features=300, n_seeds=2000,
               max_level=4, min_level=2),
    "C4": dict(w=1920, h=1080, fx=1500.0, fy=1500.0, cx=959.5, cy=539.5, n_levels=5, n_pyr=5, n_features=1000, n_seeds=10000,
               max_level=4, min_level=2),
}
CONFIGS["C1"] = CONFIGS["C2"]
CONFIGS["C5"] = CONFIGS["C2"]


# ---------------------------------------------------------------- texture
def _hash_noise(n, seed):
    """murmur3 fmix32 of the texel index -> uint8 white noise (vectorised, reproducible)."""
    i = np.arange(n, dtype=np.uint64)
    h = (i * np.uint64(0x9E3779B1) + np.uint64(seed)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    return (h >> np.uint64(24)).astype(np.int64)


def _box_sum_wrap(a, b):
    """b x b box sum with wrap-around, exact int64."""
    r = b // 2
    out = np.zeros_like(a)
    for d in range(-r, r + 1):
        out += np.roll(a, d, axis=1)
    a2 = out
    out = np.zeros_like(a)
    for d in range(-r, r + 1):
        out += np.roll(a2, d, axis=0)
    return out


def make_texture(size=2048, seed=0x00C0FFEE):
    """Multi-octave blurred noise, integer arithmetic only, tileable. uint8 [16,240]."""
    acc = np.zeros((size, size), np.int64)
    for k, (box, weight) in enumerate(((5, 3), (9, 3), (17, 2), (33, 2))):
        o = _hash_noise(size * size, seed + 7919 * k).reshape(size, size)
        o = _box_sum_wrap(_box_sum_wrap(o, box), box)
        mn, mx = int(o.min()), int(o.max())
        acc += weight * ((o - mn) * 4095 // (mx - mn))
    mn, mx = int(acc.min()), int(acc.max())
    return (16 + (acc - mn) * 224 // (mx - mn)).astype(np.uint8)


# ---------------------------------------------------------------- SE3 helpers (numpy, generation only)
def q_mul(a, b):
    x, y, z, w = a
    return np.array([w * b[0] + x * b[3] + y * b[2] - z * b[1],
                     w * b[1] + y * b[3] + z * b[0] - x * b[2],
                     w * b[2] + z * b[3] + x * b[1] - y * b[0],
                     w * b[3] - x * b[0] - y * b[1] - z * b[2]])


def q_rot(q, p):
    qv = np.asarray(q[:3])
    uv = np.cross(qv, p)
    uv = uv + uv
    return p + q[3] * uv + np.cross(qv, uv)


def q_matrix(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def se3_inverse(T):
    qi = np.array([-T[3], -T[4], -T[5], T[6]])
    return np.concatenate([-q_rot(qi, T[:3]), qi])


def se3_mul(A, B):
    return np.concatenate([A[:3] + q_rot(A[3:], B[:3]), q_mul(A[3:], B[3:])])


def se3_transform(T, p):
    return T[:3] + q_rot(T[3:], np.asarray(p, dtype=np.float64))


def se3_from_rotvec_trans(rv, t):
    rv = np.asarray(rv, dtype=np.float64)
    th = np.linalg.norm(rv)
    if th < 1e-12:
        q = np.array([0.5 * rv[0], 0.5 * rv[1], 0.5 * rv[2], 1.0])
    else:
        q = np.concatenate([np.sin(th / 2) / th * rv, [np.cos(th / 2)]])
    q = q / np.linalg.norm(q)
    return np.concatenate([np.asarray(t, dtype=np.float64), q])


def pose_error(Ta, Tb):
    """(rotation angle [rad], translation distance) between two poses."""
    E = se3_mul(Ta, se3_inverse(Tb))
    ang = 2.0 * np.arctan2(np.linalg.norm(E[3:6]), abs(E[6]))
    return ang, np.linalg.norm(E[:3])


# ---------------------------------------------------------------- trajectory + rendering
def trajectory(n_frames, seed=0x00C0FFEE, amp_scale=1.0):
    """T_f_w for n_frames: lateral sinusoid +-0.15/0.05/0.05 m, yaw/pitch +-2 deg, with seed-dependent
    phase / amplitude jitter.  Inter-frame motion is a few pixels at level 0."""
    rng = np.random.RandomState(seed & 0x7FFFFFFF)
    ph = rng.uniform(0, 2 * np.pi, 5)
    am = rng.uniform(0.8, 1.2, 5) * amp_scale
    poses = []
    for k in range(n_frames):
        s = 2 * np.pi * k / 120.0
        c = np.array([0.15 * am[0] * np.sin(s + ph[0]), 0.05 * am[1] * np.sin(1.3 * s + ph[1]),
                      0.05 * am[2] * np.sin(0.7 * s + ph[2])])
        rv = np.array([np.deg2rad(2.0) * am[3] * np.sin(0.9 * s + ph[3]), np.deg2rad(2.0) * am[4] * np.sin(1.1 * s + ph[4]), 0.0])
        T_w_f = se3_from_rotvec_trans(rv, c)
        poses.append(se3_inverse(T_w_f))
    return np.array(poses)


def render(tex, cam, T_f_w, plane_z=2.0, ppm=400.0):
    """Render the plane z=plane_z (world) textured with `tex` (ppm texels per metre, tiled)."""
    w, h, fx, fy, cx, cy = cam["w"], cam["h"], cam["fx"], cam["fy"], cam["cx"], cam["cy"]
    T_w_f = se3_inverse(np.asarray(T_f_w, dtype=np.float64))
    R = q_matrix(T_w_f[3:])
    c = T_w_f[:3]
    xs = (np.arange(w, dtype=np.float64) - cx) / fx
    ys = (np.arange(h, dtype=np.float64) - cy) / fy
    X, Y = np.meshgrid(xs, ys)
    dx = R[0, 0] * X + R[0, 1] * Y + R[0, 2]
    dy = R[1, 0] * X + R[1, 1] * Y + R[1, 2]
    dz = R[2, 0] * X + R[2, 1] * Y + R[2, 2]
    s = (plane_z - c[2]) / dz
    size = tex.shape[0]
    u = (c[0] + s * dx) * ppm + size / 2
    v = (c[1] + s * dy) * ppm + size / 2
    ui = np.floor(u).astype(np.int64)
    vi = np.floor(v).astype(np.int64)
    fu, fv = u - ui, v - vi
    u0, v0, u1, v1 = ui % size, vi % size, (ui + 1) % size, (vi + 1) % size
    t = tex.astype(np.float64)
    val = (1 - fu) * (1 - fv) * t[v0, u0] + fu * (1 - fv) * t[v0, u1] + (1 - fu) * fv * t[v1, u0] + fu * fv * t[v1, u1]
    return np.floor(val + 0.5).astype(np.uint8)


def backproject_to_plane(cam, T_f_w, px, plane_z=2.0):
    """World point on the plane seen at level-0 pixel px (ground-truth depth for features)."""
    T_w_f = se3_inverse(np.asarray(T_f_w, dtype=np.float64))
    d = q_rot(T_w_f[3:], np.array([(px[0] - cam["cx"]) / cam["fx"], (px[1] - cam["cy"]) / cam["fy"], 1.0]))
    s = (plane_z - T_w_f[2]) / d[2]
    return T_w_f[:3] + s * d
