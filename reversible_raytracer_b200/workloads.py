"""Synthetic workloads of BASELINE.json's configs, as packed tables for the
functional API (render.py).  Host-side generation with NumPy RandomState (the
seeds and distributions are SURVEY.md 8d's), results as float32 arrays.

  stress_tables   C5 / C5g: n x n image, N random spheres in the view frustum
  orbit_tables    C4: batch of planet-orbit scenes x 2 camera views
                  (orbit_experiments/planet_orbit.py:21-53, test_optimization.py:17-44)
"""
import numpy as np

from . import _native as nat


def _rotation(angle_deg, axis):
    """transform.py:95-122 (Rodrigues form, degrees, unit axis), float64."""
    a = np.asarray(axis, dtype=np.float64)
    rad = float(angle_deg) * np.pi / 180.0
    s, c = np.sin(rad), np.cos(rad)
    return np.array([
        [a[0] * a[0] + (1. - a[0] * a[0]) * c, a[0] * a[1] * (1. - c) - a[2] * s, a[0] * a[2] * (1. - c) + a[1] * s],
        [a[0] * a[1] * (1. - c) + a[2] * s, a[1] * a[1] + (1. - a[1] * a[1]) * c, a[1] * a[2] * (1. - c) - a[0] * s],
        [a[0] * a[2] * (1. - c) - a[1] * s, a[1] * a[2] * (1. - c) + a[0] * s, a[2] * a[2] + (1. - a[2] * a[2]) * c]])


def w2o_translate_scale(centres, scales):
    """w2o rows of (translate(c) * scale(s)).inverse() = scale(1/s) * translate(-c)
    (transform.py:35-38,60-93), float32 [N,12]; off-diagonals are exact zeros."""
    c = np.asarray(centres, dtype=np.float32)
    inv = (np.float32(1.0) / np.asarray(scales, dtype=np.float32)).astype(np.float32)
    N = c.shape[0]
    w = np.zeros((N, 3, 4), dtype=np.float32)
    for r in range(3):
        w[:, r, r] = inv[:, r]
        w[:, r, 3] = inv[:, r] * (-c[:, r])
    return w.reshape(N, 12)


def w2o_translate_rotate_scale(centres, angles, axes, scales):
    """w2o rows of (translate(c) * rotate(a, axis) * scale(s)).inverse()
    = scale(1/s) * rotate^T * translate(-c), float32 [N,12]."""
    N = len(centres)
    out = np.zeros((N, 3, 4), dtype=np.float32)
    for k in range(N):
        Rt = _rotation(angles[k], axes[k]).astype(np.float32).T
        Si = np.diag((np.float32(1.0) / np.asarray(scales[k], dtype=np.float32)))
        A = (Si @ Rt).astype(np.float32)
        out[k, :, :3] = A
        out[k, :, 3] = (A @ (-np.asarray(centres[k], dtype=np.float32))).astype(np.float32)
    return out.reshape(N, 12)


def stress_tables(num_objects=1024, general=False, seed=1234, centre_noise=0.0, noise_seed=1235):
    """C5 (general=False: translate*scale, diagonal A) / C5g (translate*rotate*scale).
    Returns dict(obj_type, w2o, material, light, camera) of NumPy arrays.
    centre_noise > 0 perturbs the centres with N(0, centre_noise^2) (target scenes)."""
    rng = np.random.RandomState(seed)
    N = num_objects
    z = rng.uniform(8, 16, N)
    x = rng.uniform(-0.475, 0.475, N) * z
    y = rng.uniform(-0.475, 0.475, N) * z
    if general:
        sc = rng.uniform(0.10, 0.25, (N, 3))
        ang = rng.uniform(0, 180, N)
        ax = rng.normal(size=(N, 3))
        ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    else:
        r = rng.uniform(0.10, 0.25, N)
        sc = np.stack([r, r, r], 1)
    col = rng.uniform(0.1, 1.0, (N, 3))
    ka = rng.uniform(.1, .5, N)
    kd = rng.uniform(.5, .9, N)
    centres = np.stack([x, y, z], 1)
    if centre_noise > 0:
        centres = centres + np.random.RandomState(noise_seed).normal(0, centre_noise, centres.shape)
    w2o = w2o_translate_rotate_scale(centres, ang, ax, sc) if general else w2o_translate_scale(centres, sc)
    material = np.stack([ka, kd, np.full(N, 0.3), np.full(N, 50.0), col[:, 0], col[:, 1], col[:, 2]], 1).astype(np.float32)
    light = np.array([-1., -1., 2., 0.961, 1., 0.87], dtype=np.float32)
    camera = np.concatenate([np.eye(4, dtype=np.float32)[:3].reshape(-1), [0, 0, 1]]).astype(np.float32)
    return dict(obj_type=np.zeros(N, dtype=np.int32), w2o=w2o, material=material, light=light, camera=camera,
                shader=nat.SHADER_PHONG, transpose=1)


def orbit_tables(num_scenes=256, seed=1234, centre_noise=0.0, noise_seed=1235):
    """C4: per scene Sphere(translate(centre)*scale(4)) + Sphere(translate((0,0,48))*scale(6)),
    light (0,0,1)/(1,1,1), Phong without specular, cameras translate((0,+-2.5,0)).
    Returns tables with a leading batch of 2*num_scenes (scene-major, view-minor)."""
    th = np.random.RandomState(seed).uniform(0, 2 * np.pi, num_scenes)
    centres = np.stack([9 * np.cos(th), 9 * np.sin(th), np.full(num_scenes, 32.0)], 1)
    if centre_noise > 0:
        centres = centres + np.random.RandomState(noise_seed).normal(0, centre_noise, centres.shape)
    B = 2 * num_scenes
    w2o = np.zeros((B, 2, 12), dtype=np.float32)
    camera = np.zeros((B, 15), dtype=np.float32)
    for q in range(num_scenes):
        w = w2o_translate_scale(np.array([centres[q], [0, 0, 48]]), np.array([[4, 4, 4], [6, 6, 6]]))
        for v in range(2):
            w2o[2 * q + v] = w
            cam = np.eye(4, dtype=np.float32)
            cam[1, 3] = 2.5 if v == 0 else -2.5
            camera[2 * q + v] = np.concatenate([cam[:3].reshape(-1), [0, 0, 1]])
    material = np.array([[0.5, 0.7, 0.3, 50., 0.0, 0.9, 0.0], [0.4, 0.9, 0.3, 50., 0.9, 0.0, 0.0]], dtype=np.float32)
    light = np.array([0., 0., 1., 1., 1., 1.], dtype=np.float32)
    return dict(obj_type=np.zeros(2, dtype=np.int32), w2o=w2o, material=material, light=light, camera=camera,
                shader=nat.SHADER_PHONG_NOSPEC, transpose=0, centres=centres)


def algorithmic_flops(num_rays, num_objects, hit_rays, general=False, backward=True):
    """16 flops per diagonal ray-sphere test (28 general), one sweep
    credited, + 64 (shade) and 256 (reverse pass) nominal flops per winning ray."""
    f_test = 28 if general else 16
    return float(num_rays) * num_objects * f_test + float(hit_rays) * (64 + (256 if backward else 0))
