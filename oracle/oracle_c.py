"""oracle_c -- TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/oracle_c.c.

Never imported by the product package.  Builds oracle/_ref/liboracle_c.so with
`make -C oracle` when it is missing (gcc is in the image).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, '_ref', 'liboracle_c.so')

SHADERS = {'phong': 0, 'phong_nospec': 1, 'depth': 2}


class RrtScene(C.Structure):
    """Mirror of `struct rrt_scene` (include/rrt_b200.h)."""
    _fields_ = [
        ('n', C.c_int32), ('samples', C.c_int32), ('num_objects', C.c_int32),
        ('num_scenes', C.c_int32), ('shader', C.c_int32), ('transpose', C.c_int32),
        ('row_begin', C.c_int32), ('row_count', C.c_int32), ('max_depth', C.c_float),
        ('camera_grad', C.c_int32), ('seed', C.c_uint64),
        ('obj_type', C.c_void_p), ('w2o', C.c_void_p), ('material', C.c_void_p),
        ('light', C.c_void_p), ('camera', C.c_void_p),
        ('jitter_x', C.c_void_p), ('jitter_y', C.c_void_p),
        ('w2o_scene_stride', C.c_int64), ('material_scene_stride', C.c_int64),
        ('light_scene_stride', C.c_int64), ('camera_scene_stride', C.c_int64),
        ('jitter_scene_stride', C.c_int64), ('base_rays', C.c_void_p),
        ('scene_begin', C.c_int32), ('flags', C.c_int32), ('obj_records', C.c_void_p),
        ('ticket', C.c_void_p), ('det_workspace', C.c_void_p),
        ('reflectivity', C.c_void_p), ('reflectivity_scene_stride', C.c_int64),
    ]


def build(force=False):
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, 'oracle_c.c')):
        subprocess.check_call(['make', '-C', HERE, '-s', '-B'] if force else ['make', '-C', HERE, '-s'],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_rng_value.restype = C.c_float
        _lib.orc_rng_value.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    return _lib


def num_threads():
    return int(lib().orc_num_threads())


def use_all_cores():
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline legs want every core this
    process may run on."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    lib().orc_set_num_threads(int(n))
    return num_threads()


FLAG_SHADOWS, HIT_SHADOWED, FLAG_MIRROR = 4, 0x40000000, 128


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class PackedScene:
    """Flat float32 tables in the layout of include/rrt_b200.h, from a spec dict
    (oracle_numpy.py) or from already-packed arrays.  Keeps the arrays alive."""

    def __init__(self, n, samples, obj_type, w2o, material, light, camera, shader,
                 transpose, max_depth=1.0, jitter_x=None, jitter_y=None, seed=0,
                 camera_grad=0, row_begin=0, row_count=0, scene_begin=0, shadows=0, reflectivity=None):
        self.n, self.samples = int(n), int(samples)
        self.obj_type = np.ascontiguousarray(obj_type, dtype=np.int32)
        self.N = int(self.obj_type.shape[0])
        w2o = np.ascontiguousarray(w2o, dtype=np.float32)
        self.B = 1 if w2o.ndim == 2 else int(w2o.shape[0])
        self.w2o = w2o.reshape(self.B, self.N, 12)
        self.material = np.ascontiguousarray(material, dtype=np.float32).reshape(-1, self.N, 7)
        self.light = np.ascontiguousarray(light, dtype=np.float32).reshape(-1, 6)
        self.camera = np.ascontiguousarray(camera, dtype=np.float32).reshape(-1, 15)
        self.shader = SHADERS[shader] if isinstance(shader, str) else int(shader)
        self.transpose = int(transpose)
        self.max_depth = float(max_depth)
        self.seed = int(seed)
        self.camera_grad = int(camera_grad)
        self.row_begin = int(row_begin)
        self.row_count = int(row_count)
        self.scene_begin = int(scene_begin)
        self.shadows = int(shadows)
        # mirror bounce (RRT_FLAG_MIRROR): per-object reflectivity [N] (or [B,N]); None = no bounce
        self.reflectivity = None if reflectivity is None else np.ascontiguousarray(reflectivity, dtype=np.float32)
        self.rows = self.row_count if self.row_count > 0 else self.n - self.row_begin
        self.jitter_x = None if jitter_x is None else np.ascontiguousarray(jitter_x, dtype=np.float32)
        self.jitter_y = None if jitter_y is None else np.ascontiguousarray(jitter_y, dtype=np.float32)

    @staticmethod
    def from_spec(spec, camera_grad=None, use_rng_seed=None):
        """spec dict (oracle_numpy.py) -> packed tables.  Jitter is converted from
        the reference's ray index space to IMAGE index space."""
        root = spec.get('cam_o2w') is None
        cam = np.eye(4, dtype=np.float32) if root else np.asarray(spec['cam_o2w'], dtype=np.float32)
        camera = np.concatenate([cam[:3, :].reshape(-1), np.asarray(spec['look_at'], dtype=np.float32)])
        jx, jy = spec.get('jitter_x'), spec.get('jitter_y')
        if use_rng_seed is not None:
            jx = jy = None
        elif root:
            jx, jy = jx.transpose(1, 0, 2), jy.transpose(1, 0, 2)
        return PackedScene(
            spec['n'], spec['samples'], spec['obj_type'],
            np.asarray(spec['w2o'], dtype=np.float32)[:, :3, :].reshape(-1, 12),
            spec['material'],
            np.concatenate([spec['light_dir'], spec['light_int']]), camera, spec['shader'],
            transpose=1 if root else 0, max_depth=spec.get('max_depth', 1.0),
            jitter_x=jx, jitter_y=jy, seed=use_rng_seed or 0,
            camera_grad=(0 if root else 1) if camera_grad is None else camera_grad,
            shadows=int(bool(spec.get('shadows', 0))), reflectivity=spec.get('reflectivity'))

    def slab(self, row_begin, row_count):
        jx = None if self.jitter_x is None else self.jitter_x.reshape(-1, self.n, self.n, self.samples)[:, row_begin:row_begin + row_count]
        jy = None if self.jitter_y is None else self.jitter_y.reshape(-1, self.n, self.n, self.samples)[:, row_begin:row_begin + row_count]
        return PackedScene(self.n, self.samples, self.obj_type, self.w2o, self.material, self.light,
                           self.camera, self.shader, self.transpose, self.max_depth, jx, jy, self.seed,
                           self.camera_grad, row_begin, row_count, shadows=self.shadows, reflectivity=self.reflectivity)

    def desc(self):
        def stride(a, per):
            return 0 if a.shape[0] == 1 else per
        d = RrtScene()
        d.n, d.samples, d.num_objects, d.num_scenes = self.n, self.samples, self.N, self.B
        d.shader, d.transpose = self.shader, self.transpose
        d.row_begin, d.row_count = self.row_begin, self.row_count
        d.scene_begin = self.scene_begin
        d.flags = (FLAG_SHADOWS if self.shadows else 0) | (FLAG_MIRROR if self.reflectivity is not None else 0)
        if self.reflectivity is not None:
            d.reflectivity = _ptr(self.reflectivity)
            d.reflectivity_scene_stride = 0 if self.reflectivity.size == self.N else self.N
        d.max_depth, d.camera_grad, d.seed = self.max_depth, self.camera_grad, self.seed
        d.obj_type, d.w2o, d.material = _ptr(self.obj_type), _ptr(self.w2o), _ptr(self.material)
        d.light, d.camera = _ptr(self.light), _ptr(self.camera)
        d.jitter_x, d.jitter_y = _ptr(self.jitter_x), _ptr(self.jitter_y)
        d.w2o_scene_stride = stride(self.w2o, self.N * 12)
        d.material_scene_stride = stride(self.material, self.N * 7)
        d.light_scene_stride = stride(self.light, 6)
        d.camera_scene_stride = stride(self.camera, 15)
        if self.jitter_x is not None:
            per = self.rows * self.n * self.samples
            d.jitter_scene_stride = 0 if self.jitter_x.size == per else per
        return d

    @property
    def grad_size(self):
        return self.N * 19 + 21


def render_forward(ps, want_aux=True):
    d = ps.desc()
    image = np.zeros((ps.B, ps.rows, ps.n, 3), dtype=np.float32)
    hit = np.zeros((ps.B, ps.samples, ps.rows, ps.n), dtype=np.int32) if want_aux else None
    tmin = np.zeros((ps.B, ps.samples, ps.rows, ps.n), dtype=np.float32) if want_aux else None
    rc = lib().orc_render_forward(C.byref(d), _ptr(image), _ptr(hit), _ptr(tmin))
    assert rc == 0, rc
    return image, hit, tmin


def render_forward_secondary(ps):
    """-> image, hit_index, secondary (mirror) hit index [B,S,rows,n] (-1 = none)"""
    d = ps.desc()
    image = np.zeros((ps.B, ps.rows, ps.n, 3), dtype=np.float32)
    hit = np.zeros((ps.B, ps.samples, ps.rows, ps.n), dtype=np.int32)
    hit2 = np.zeros((ps.B, ps.samples, ps.rows, ps.n), dtype=np.int32)
    rc = lib().orc_render_forward_secondary(C.byref(d), _ptr(image), _ptr(hit), _ptr(hit2))
    assert rc == 0, rc
    return image, hit, hit2


def render_backward(ps, dl_dimage, hit_index=None):
    d = ps.desc()
    g = np.ascontiguousarray(dl_dimage, dtype=np.float32).reshape(ps.B, ps.rows, ps.n, 3)
    h = None if hit_index is None else np.ascontiguousarray(hit_index, dtype=np.int32)
    grad = np.zeros((ps.B, ps.grad_size), dtype=np.float64)
    rc = lib().orc_render_backward(C.byref(d), _ptr(g), _ptr(h), _ptr(grad))
    assert rc == 0, rc
    return grad


def render_fused_mse(ps, target, channel_weight=None):
    d = ps.desc()
    t = np.ascontiguousarray(target, dtype=np.float32).reshape(ps.B, ps.rows, ps.n, 3)
    cw = None if channel_weight is None else np.ascontiguousarray(channel_weight, dtype=np.float32)
    image = np.zeros((ps.B, ps.rows, ps.n, 3), dtype=np.float32)
    hit = np.zeros((ps.B, ps.samples, ps.rows, ps.n), dtype=np.int32)
    loss = np.zeros(ps.B, dtype=np.float64)
    grad = np.zeros((ps.B, ps.grad_size), dtype=np.float64)
    rc = lib().orc_render_fused_mse(C.byref(d), _ptr(t), _ptr(cw), _ptr(image), _ptr(hit), _ptr(loss), _ptr(grad))
    assert rc == 0, rc
    return image, hit, loss, grad


def primary_rays(ps):
    d = ps.desc()
    out = np.zeros((ps.rows, ps.n, ps.samples, 3), dtype=np.float32)
    rc = lib().orc_primary_rays(C.byref(d), _ptr(out))
    assert rc == 0, rc
    return out


def rng_value(seed, scene, pix, s, axis):
    return float(lib().orc_rng_value(seed, scene, pix, s, axis))


def split_grad(flat, N):
    """flat[N*19+21] -> dict in the layout of include/rrt_b200.h"""
    flat = np.asarray(flat)
    og = flat[:N * 19].reshape(N, 19)
    gg = flat[N * 19:]
    return dict(w2o=og[:, :12].reshape(N, 3, 4), material=og[:, 12:19],
                light_dir=gg[0:3], light_int=gg[3:6], cam_o2w=gg[6:18].reshape(3, 4), look_at=gg[18:21])
