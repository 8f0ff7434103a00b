"""Shared test helpers: oracle PackedScene -> device tensors for the product API."""
import numpy as np
import torch

from reversible_raytracer_b200.render import RenderConfig

# Set by the parity module's `kernel_choice` fixture: True forces the general 8-rays-per-thread
# kernel where the small-scene kernel would be chosen, so every case is checked on both.
NO_SMALL = False
PIXEL_THREADS = 0      # 1: force the pixel-per-thread mapping of the small-scene kernel where it applies, 2: ray-per-thread
USE_RECORDS = True     # False: the kernels build the sweep records per CTA instead of TMA-staging a prebuilt table


def to_device(ps, device, with_jitter=True):
    """oracle.oracle_c.PackedScene -> (cfg, obj_type, w2o, material, light, camera, jitter)."""
    cfg = RenderConfig(n=ps.n, samples=ps.samples, shader=ps.shader, transpose=ps.transpose,
                       max_depth=ps.max_depth, camera_grad=ps.camera_grad, seed=ps.seed,
                       row_begin=ps.row_begin, row_count=ps.row_count, scene_begin=getattr(ps, 'scene_begin', 0),
                       no_small=int(NO_SMALL), shadows=int(getattr(ps, 'shadows', 0)),
                       use_records=int(USE_RECORDS), pixel_threads=int(PIXEL_THREADS))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    w2o = t(ps.w2o) if ps.B > 1 else t(ps.w2o[0])
    jitter = None
    if with_jitter and ps.jitter_x is not None:
        jitter = (t(ps.jitter_x), t(ps.jitter_y))
    return cfg, t(ps.obj_type), w2o, t(ps.material), t(ps.light), t(ps.camera), jitter


def reflectivity_of(ps, device):
    """mirror bounce: the PackedScene's per-object reflectivity as a device tensor, or None"""
    r = getattr(ps, 'reflectivity', None)
    return None if r is None else torch.from_numpy(np.ascontiguousarray(r)).to(device)


def block_rel_err(a, b):
    """max |a-b| / max |b| over one parameter block (gradient tolerance metric)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    s = np.max(np.abs(b)) if b.size else 0.0
    if s == 0.0:
        return float(np.max(np.abs(a))) if a.size else 0.0
    return float(np.max(np.abs(a - b)) / s)
