"""Scene / Camera / Light / Material -- host-side mirror of the reference's
scene.py (root variant) and orbit_experiments/scene.py (camera with a transform).

    Scene(shapes, lights, camera, shader).build(antialias_samples=4)   scene.py:11-52
    Camera(x_dims, y_dims)                       root variant          scene.py:55-75
    Camera(x_dims, y_dims, o2w, camera_dir)      orbit variant         orbit_experiments/scene.py:55-80
    Light(direction, intensity).normed_dir()                           scene.py:78-86
    Material(color, ks, kd, ka, shininess)                             scene.py:89-101

Execution-model bridge: the reference's build() returns a SYMBOLIC
image that is compiled once and re-evaluated after its theano.shared parameters
change.  Here parameters are torch tensors (requires_grad=True for the ones being
optimised) and build() renders EAGERLY with the CUDA kernels, returning an image
tensor wired into torch autograd.  Call build() again after updating parameters.
Like the reference's compiled graph, a Scene reuses the SAME anti-alias jitter on
every build() (scene.py:24-25 bakes it into the graph as constants); it is drawn
from NumPy's global RNG on first use unless `jitter=`/`seed=` is given.

There is no CPU fallback: build() raises if CUDA or the extension is missing.
"""
import numpy as np
import torch

from . import _native as nat
from . import render as R
from .transform import RayField, Transform, _Arg, as_tensor, default_device, identity  # noqa: F401
from .shape import *  # noqa: F401,F403  (the reference's scene.py re-exports these)
from .util import *  # noqa: F401,F403
from .transform import *  # noqa: F401,F403


# Module-level caches of host-generated constants, so that a loss closure which rebuilds its
# Scene on every call (orbit_experiments/test_optimization.py:17-44 does) performs no
# host->device copy in steady state and can be captured into a CUDA graph.
_SEEDED_JITTER = {}      # (n, S, seed, transposed, device) -> (jx, jy) device tensors
_OBJ_TYPES = {}          # (kinds, device) -> int32 device tensor


def _obj_type_tensor(kinds, device):
    key = (tuple(int(k) for k in kinds), str(device))
    t = _OBJ_TYPES.get(key)
    if t is None:
        if len(_OBJ_TYPES) > 1024:
            _OBJ_TYPES.clear()
        t = _OBJ_TYPES[key] = torch.tensor(list(key[0]), dtype=torch.int32, device=device)
    return t


def _live_field(name):
    """A Material / Light / Camera field that behaves like a member holding a theano.shared or a symbolic
    slice of one: assigning stores a transform._Arg (constants: by-value cached device tensor; tensors:
    live, cast to float32 at read time; VIEWS OF A LEAF such as `q[:3]` are re-taken on every read, so they
    follow a later requires_grad_ and in-place updates and pin no autograd node to the construction site),
    reading returns the current float32 tensor."""
    slot = '_' + name

    def get(self):
        return getattr(self, slot).t

    def put(self, value):
        setattr(self, slot, _Arg(value))
    return property(get, put)


def _field_key(obj, names):
    """identity of the values behind live fields, for Scene's cache signature"""
    out = []
    for n in names:
        a = getattr(obj, '_' + n)
        out.append(a.key if a.key is not None else ('const', id(a._src)))
    return tuple(out)


class Material(object):
    """scene.py:89-101"""
    ks, kd, ka = _live_field('ks'), _live_field('kd'), _live_field('ka')
    color, shininess = _live_field('color'), _live_field('shininess')
    FIELDS = ('ka', 'kd', 'ks', 'shininess', 'color')

    def __init__(self, color, ks, kd, ka, shininess, reflectivity=0.0):
        """`reflectivity` (extension, default 0 = the reference's behaviour): mirror coefficient k in
        [0,1] of one reflection bounce, rgb = (1-k) rgb + k rgb_seen_along_the_reflected_ray
        (RRT_FLAG_MIRROR in include/rrt_b200.h; a constant, root camera variant only)."""
        self.reflectivity = float(reflectivity)
        # live parameters (torch tensors) are re-read on every build; constants are packed once
        self.dynamic = any(isinstance(v, torch.Tensor) for v in (color, ks, kd, ka, shininess))
        self.ks, self.kd, self.ka, self.color, self.shininess = ks, kd, ka, color, shininess

    def packed(self, device):
        """-> float32[7] (ka, kd, ks, shininess, r, g, b), differentiable."""
        parts = [self.ka.reshape(1), self.kd.reshape(1), self.ks.reshape(1), self.shininess.reshape(1),
                 self.color.reshape(3)]
        return torch.cat([p.to(device) for p in parts])


class Light(object):
    """Directional light, scene.py:78-86."""
    direction, intensity = _live_field('direction'), _live_field('intensity')
    FIELDS = ('direction', 'intensity')

    def __init__(self, direction, intensity):
        self.dynamic = isinstance(direction, torch.Tensor) or isinstance(intensity, torch.Tensor)
        self.direction, self.intensity = direction, intensity

    def normed_dir(self):
        d = self.direction
        norm = torch.sqrt(d[0] ** 2 + d[1] ** 2 + d[2] ** 2)
        return d / norm

    def packed(self, device):
        return torch.cat([self.direction.reshape(3).to(device), self.intensity.reshape(3).to(device)])


class Camera(object):
    """Pin-hole camera.  Two constructors, like the reference's two copies."""
    look_at = _live_field('look_at')

    def __init__(self, x_dims, y_dims, o2w=None, camera_dir=None):
        self.x_dims = int(x_dims)
        self.y_dims = int(y_dims)
        self.has_transform = o2w is not None
        self.o2w = o2w if o2w is not None else identity()
        self.w2o = self.o2w.inverse()
        self.look_at = np.asarray([0, 0, 1.], dtype='float32') if camera_dir is None else camera_dir
        self.dynamic_look_at = isinstance(camera_dir, torch.Tensor)
        self._rays = None
        self._last_sample = None     # (sampleDist_x, sampleDist_y) of the last build's last sample

    @property
    def rays(self):
        """The reference leaves the LAST anti-alias sample's RayField on the camera after
        build() (scene.py:30-32); some scripts read it (e.g. autoencoder_2ly.py:139).  The
        kernels never materialise rays, so this dense field is rebuilt on demand."""
        if self._rays is None and self._last_sample is not None:
            _, jit, S, transposed = self._last_sample
            jx, jy = (j[:, :, S - 1].detach().cpu().numpy() for j in jit)      # image index space
            if transposed:
                jx, jy = jx.T, jy.T                                            # back to ray index space
            sx = (jx + np.float32(S - 1)) / np.float32(S)                      # scene.py:31-32
            sy = (jy + np.float32(S - 1)) / np.float32(S)
            self._rays = self.make_rays(self.x_dims, self.y_dims, sx, sy)
        return self._rays

    @rays.setter
    def rays(self, value):
        self._rays = value

    def make_rays(self, x_dims, y_dims, sampleDist_x=None, sampleDist_y=None):
        """scene.py:61-75 (dense helper; the kernels generate rays in registers with
        the same float64 -> float32 arithmetic).  Orbit variant applies camera.o2w
        (orbit_experiments/scene.py:80)."""
        rays = np.dstack(np.meshgrid(np.linspace(0.5, -0.5, y_dims),
                                     np.linspace(-0.5, 0.5, x_dims), indexing='ij'))
        rays = np.dstack([rays, np.ones([y_dims, x_dims], dtype='float32')])
        rays = np.divide(rays, np.linalg.norm(rays, axis=2).reshape(y_dims, x_dims, 1).repeat(3, 2))
        rays = np.asarray(rays, dtype=np.float32)
        if sampleDist_x is not None:
            rays[:, :, 0] = rays[:, :, 0] + np.asarray(sampleDist_x, dtype=np.float32) / np.float32(x_dims)
        if sampleDist_y is not None:
            rays[:, :, 1] = rays[:, :, 1] + np.asarray(sampleDist_y, dtype=np.float32) / np.float32(y_dims)
        rf = RayField([0., 0., 0.], rays)
        return self.o2w(rf) if self.has_transform else rf

    def packed(self, device):
        """-> float32[15]: camera.o2w rows 0..2 (3x4) then look_at."""
        return torch.cat([self.o2w.m[:3, :].reshape(12).to(device), self.look_at.reshape(3).to(device)])


class Scene(object):
    """scene.py:11-52"""

    def __init__(self, shapes, lights, camera, shader, shadows=False, deterministic=False):
        """`shadows=True` switches on the hard-shadow pass the reference has commented out
        (scene.py:41-45, Sphere.shadow shape.py:85-97; semantics in include/rrt_b200.h) --
        an extension, off by default like in the reference.  `deterministic=True` makes gradients
        and losses bit-identical from run to run, like the reference's T.grad (optimize.py:25):
        RRT_FLAG_DETERMINISTIC, fixed-point accumulation instead of float atomics."""
        self.shadows = bool(shadows)
        self.deterministic = bool(deterministic)
        self.shapes = shapes
        self.lights = lights
        self.camera = camera
        self.shader = shader
        self._jitter = {}
        self._cache = None
        self.last_hit_index = None

    # -- jitter ---------------------------------------------------------------
    def _jitter_for(self, n, S, jitter, seed, device):
        """Anti-alias offsets in IMAGE index space (include/rrt_b200.h).  The reference
        draws x then y as (x_dims, y_dims, S) arrays from the global RNG
        (scene.py:24-25) in RAY index space; the root variant shades pixel (a,b) with
        ray [b,a], hence the transpose."""
        key = (n, S)
        transposed = not self.camera.has_transform
        if jitter is not None:
            jx, jy = (np.asarray(j, dtype=np.float32) for j in jitter)
        elif seed is not None:
            # seeded draws are reproducible: generated and uploaded once per process (module-level
            # cache, so even a Scene rebuilt inside a loss closure stays free of host->device copies)
            gkey = (n, S, int(seed), transposed, str(device))
            out = _SEEDED_JITTER.get(gkey)
            if out is None:
                rng = np.random.RandomState(seed)
                jx = np.asarray(rng.random_sample((n, n, S)), dtype=np.float32)
                jy = np.asarray(rng.random_sample((n, n, S)), dtype=np.float32)
                if transposed:
                    jx, jy = jx.transpose(1, 0, 2), jy.transpose(1, 0, 2)
                if len(_SEEDED_JITTER) > 64:
                    _SEEDED_JITTER.clear()
                out = _SEEDED_JITTER[gkey] = (torch.from_numpy(np.ascontiguousarray(jx)).to(device),
                                              torch.from_numpy(np.ascontiguousarray(jy)).to(device))
            self._jitter[key] = out
            return out
        elif key in self._jitter:
            return self._jitter[key]
        else:
            jx = np.asarray(np.random.random((n, n, S)), dtype=np.float32)
            jy = np.asarray(np.random.random((n, n, S)), dtype=np.float32)
        if transposed:
            jx, jy = jx.transpose(1, 0, 2), jy.transpose(1, 0, 2)
        out = (torch.from_numpy(np.ascontiguousarray(jx)).to(device),
               torch.from_numpy(np.ascontiguousarray(jy)).to(device))
        self._jitter[key] = out
        return out

    def reset_jitter(self):
        self._jitter = {}

    # -- packing ----------------------------------------------------------------
    def device(self):
        if self._cache is not None:
            return self._cache['device']
        d = default_device()
        if d.type == 'cuda':
            return d
        for s in self.shapes:
            if s.w2o.m.is_cuda:
                return s.w2o.m.device
        return d

    def _static(self, device):
        """Structure-dependent state cached across builds: the compiled transform chain
        (chain.py), obj_type, and the packed tables of constant materials / light / camera."""
        st = self._cache
        shapes = list(self.shapes)
        light = self.lights[0]
        # everything the cached tables were packed from, by identity: shapes, their transforms and
        # materials (and the constant materials' field tensors), the light's fields, the camera
        sig = (device, tuple(id(s) for s in shapes), tuple(id(s.w2o) for s in shapes),
               tuple((id(s.material),) + _field_key(s.material, Material.FIELDS) for s in shapes),
               id(light), _field_key(light, Light.FIELDS),
               id(self.camera), id(self.camera.o2w), _field_key(self.camera, ('look_at',)))
        if st is not None and st['sig'] == sig:
            return st
        from .chain import ChainProgram
        # (the objects themselves are kept alive in the cache entry, so the ids in `sig` stay unique)
        st = dict(sig=sig, device=device, shapes=shapes, w2o=[s.w2o for s in shapes], camera=self.camera, light=light,
                  materials=[s.material for s in shapes])
        st['obj_type'] = _obj_type_tensor([s.kind for s in shapes], device)
        # two programs, so that the shapes' rows ARE the renderer's w2o table (no slicing of a
        # joint output and no scatter of its gradient), and a constant camera is evaluated once
        st['prog'] = st['cam_prog'] = None
        if device.type == 'cuda':
            try:
                st['prog'] = ChainProgram([s.w2o for s in shapes], device) if shapes else None
            except ValueError:
                pass                               # explicit-matrix transforms: torch path
            try:
                st['cam_prog'] = ChainProgram([self.camera.o2w], device)
            except ValueError:
                pass
        st['cam_t'] = None
        st['mat'] = None
        if shapes and not any(s.material.dynamic for s in shapes):
            st['mat'] = torch.stack([s.material.packed(device) for s in shapes]).detach()
        st['light_t'] = None if self.lights[0].dynamic else self.lights[0].packed(device).detach()
        refl = [float(getattr(s.material, 'reflectivity', 0.0)) for s in shapes]
        st['refl'] = as_tensor(np.asarray(refl, dtype=np.float32), device=device) if any(r != 0.0 for r in refl) else None
        self._cache = st
        return st

    def pack(self, device=None):
        """-> (obj_type int32[N], w2o [N,12], material [N,7], light [6], camera [15]),
        all on `device`, differentiable w.r.t. whatever the user's tensors require."""
        device = device or self.device()
        st = self._static(device)
        N = len(self.shapes)
        if st['prog'] is not None:
            w2o = st['prog'].evaluate()            # one kernel: every shape's w2o rows
        else:
            w2o = torch.stack([s.w2o.m[:3, :].reshape(12).to(device) for s in self.shapes]) if N else \
                torch.zeros((0, 12), dtype=torch.float32, device=device)
        if st['mat'] is not None:
            mat = st['mat']
        elif N:
            mat = torch.stack([s.material.packed(device) for s in self.shapes])
        else:
            mat = torch.zeros((0, 7), dtype=torch.float32, device=device)
        light = st['light_t'] if st['light_t'] is not None else self.lights[0].packed(device)
        look = self.camera.look_at
        cam_static = st['cam_prog'] is not None and not st['cam_prog'].dynamic and not look.requires_grad
        if cam_static and st['cam_t'] is not None:
            cam = st['cam_t']                      # constant camera: packed once
        else:
            cam_rows = st['cam_prog'].evaluate()[0] if st['cam_prog'] is not None else \
                self.camera.o2w.m[:3, :].reshape(12).to(device)
            cam = torch.cat([cam_rows, look.reshape(3).to(device)])
            if cam_static:
                st['cam_t'] = cam = cam.detach()
        return st['obj_type'], w2o, mat, light, cam

    CULL_MIN_OBJECTS = 32    # Scene.build turns conservative culling on from this many shapes

    def config(self, antialias_samples=4, cull=None):
        cam = self.camera
        if cam.x_dims != cam.y_dims:
            raise ValueError('the reference renderer only works for x_dims == y_dims '
                             '(rays are (y,x,3), image and jitter are (x,y,.): scene.py:21,24)')
        return R.RenderConfig(n=cam.x_dims, samples=int(antialias_samples), shader=self.shader.shader_id,
                              transpose=0 if cam.has_transform else 1,
                              max_depth=float(getattr(self.shader, 'maxDepth', 1.0)),
                              camera_grad=1 if cam.has_transform else 0,
                              cull=int(len(self.shapes) >= self.CULL_MIN_OBJECTS if cull is None else bool(cull)),
                              shadows=int(self.shadows), geom_grad_only=int(self._geom_grad_only()),
                              deterministic=int(self.deterministic))

    def _geom_grad_only(self):
        """True when no gradient can be asked for materials, light or look_at -- they were all given
        as constants (tuples / NumPy), as in every decoder of the reference (autoencoder.py:57-71,
        orbit_experiments/test_optimization.py:17-44).  The reverse pass then skips those sums
        (RRT_FLAG_NO_MATERIAL_GRAD)."""
        look = self.camera.look_at
        return (not any(s.material.dynamic for s in self.shapes) and not self.lights[0].dynamic
                and not self.camera.dynamic_look_at and not look.requires_grad)

    # -- rendering ------------------------------------------------------------------
    def _prepare(self, antialias_samples, jitter, seed, cull):
        """Shared prologue of build() / build_mse(): device, config, packed tables, jitter and the
        camera's last-sample bookkeeping (scene.py:30-32 leaves the last RayField on the camera)."""
        if not torch.cuda.is_available():
            raise nat.NativeError('Scene.build needs a CUDA device (B200); there is no CPU fallback')
        device = self.device()
        if device.type != 'cuda':
            device = torch.device('cuda', torch.cuda.current_device())
        cfg = self.config(antialias_samples, cull)     # culling never changes a bit of the result
        tables = self.pack(device)
        jit = self._jitter_for(cfg.n, cfg.samples, jitter, seed, device)
        self.camera._rays, self.camera._last_sample = None, ('lazy', jit, cfg.samples, not self.camera.has_transform)
        self._refl = self._static(device)['refl']          # mirror bounce (Material.reflectivity), or None
        return device, cfg, tables, jit

    def build(self, antialias_samples=4, jitter=None, seed=None, cull=None):
        """Render the scene (scene.py:18-52) -> image (x_dims, y_dims, 3) float32 on the
        GPU, differentiable w.r.t. shape transforms, materials, the light and (orbit
        variant) the camera transform."""
        device, cfg, (obj_type, w2o, mat, light, cam), jit = self._prepare(antialias_samples, jitter, seed, cull)
        return R.render(cfg, obj_type, w2o, mat, light, cam, jit, self._refl)

    def build_mse(self, target, antialias_samples=4, channel_weight=None, jitter=None, seed=None,
                  want_image=False, cull=None, linear=False):
        """Fused forward + sum((image-target)^2) + reverse pass in ONE kernel (the cost of
        match_mirror.py:45 and of every autoencoder, autoencoder.py:76).  Returns a
        differentiable scalar loss (float32) -- call .backward() on it -- and, if asked,
        the detached image."""
        device, cfg, (obj_type, w2o, mat, light, cam), jit = self._prepare(antialias_samples, jitter, seed, cull)
        if linear:                                  # RRT_FLAG_LINEAR_COST: `target` is the weight image W
            from dataclasses import replace
            cfg = replace(cfg, linear_cost=1)
        loss, image = _FusedMSE.apply(w2o, mat, light, cam, cfg, obj_type, jit,
                                      as_tensor(target).to(device), channel_weight, want_image, self._refl)
        return (loss, image) if want_image else loss

    def build_linear(self, weights, antialias_samples=4, channel_weight=None, jitter=None, seed=None,
                     want_image=False, cull=None):
        """Fused forward + LINEAR cost sum(weights * image) + reverse pass in one kernel: the loss of
        optimize_brightness.py:51, `-image[90,85].sum() - image[50,90].sum()`, is a weight image that is
        -1 at two pixels and 0 elsewhere (rays of zero-weight pixels skip the reverse pass).  Returns a
        differentiable scalar like build_mse."""
        return self.build_mse(weights, antialias_samples, channel_weight, jitter, seed, want_image, cull, linear=True)

    def linear_cost(self, weights, antialias_samples=4, channel_weight=None, jitter=None, seed=None):
        """The cost expression `(weights * scene.build()).sum()` as a closure for GDOptimizer.optimize --
        like mse_cost: fused kernel, and the whole optimise step as ONE launch where the scene qualifies."""
        def cost():
            return self.build_linear(weights, antialias_samples, channel_weight, jitter, seed)
        cost.fused_spec = dict(scene=self, target=weights, antialias_samples=antialias_samples,
                               channel_weight=channel_weight, jitter=jitter, seed=seed, linear=True)
        return cost

    def mse_cost(self, target, antialias_samples=4, channel_weight=None, jitter=None, seed=None):
        """The cost expression `((scene.build() - target) ** 2).sum()` (match_mirror.py:45) as a
        closure for GDOptimizer.optimize: calling it renders through the fused kernel like
        build_mse; GDOptimizer additionally recognises it and, when the optimised variables
        are exactly the parameters of the shapes' transforms of a small scene, runs the WHOLE
        step (chains, render, loss, reverse pass, update) as one kernel launch
        (rrt_small_step_mse)."""
        def cost():
            return self.build_mse(target, antialias_samples, channel_weight, jitter, seed)
        cost.fused_spec = dict(scene=self, target=target, antialias_samples=antialias_samples,
                               channel_weight=channel_weight, jitter=jitter, seed=seed)
        return cost


class _FusedMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w2o, mat, light, cam, cfg, obj_type, jit, target, channel_weight, want_image, reflectivity=None):
        loss, grad, image, _ = R.render_fused_mse(cfg, obj_type, w2o, mat, light, cam, target, channel_weight, jit,
                                                  want_image=want_image, reflectivity=reflectivity)
        ctx.N = w2o.shape[-2]
        ctx.save_for_backward(grad)
        if image is None:
            image = torch.empty(0, device=w2o.device)
        ctx.mark_non_differentiable(image)
        return loss.float(), image

    @staticmethod
    def backward(ctx, g_loss, _g_image):
        (grad,) = ctx.saved_tensors
        gw, gm, gl, gc = R.split_grad(grad * g_loss, ctx.N)
        return gw, gm, gl, gc, None, None, None, None, None, None, None
