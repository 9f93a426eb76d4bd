"""Shared synthetic scenes for the tests (deterministic, seeded)."""
import functools
import numpy as np
from android_svo_b200 import synth


@functools.lru_cache(maxsize=None)
def texture(size=1024):
    return synth.make_texture(size)


@functools.lru_cache(maxsize=None)
def scene(name="C2", n_frames=10, stride=4, amp=1.5, tex_size=1024):
    """n_frames views of the plane, every `stride`-th pose of the trajectory."""
    cfg = synth.CONFIGS[name]
    poses = synth.trajectory(n_frames * stride, amp_scale=amp)[::stride]
    imgs = [synth.render(texture(tex_size), cfg, T) for T in poses]
    return cfg, poses, imgs


def cam_of(cfg, cls):
    return cls.make(cfg["w"], cfg["h"], cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"])


def noise_image(h, w, seed, blur=True):
    rng = np.random.RandomState(seed)
    img = rng.randint(0, 256, (h, w)).astype(np.uint8)
    if blur:  # 3x3 box blur keeps FAST busy without saturating it
        f = img.astype(np.int32)
        acc = np.zeros_like(f)
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                acc += np.roll(np.roll(f, dy, 0), dx, 1)
        img = (acc // 9).astype(np.uint8)
    return img


def gt_depth(cfg, T_f_w, px):
    p = synth.backproject_to_plane(cfg, T_f_w, px)
    return np.linalg.norm(p - synth.se3_inverse(T_f_w)[:3]), p
