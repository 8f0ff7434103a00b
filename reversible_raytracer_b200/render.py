"""Functional layer over the C ABI: torch tensors in, torch tensors out.

PyTorch is used for device buffers, streams and autograd plumbing only; every
number is produced by the kernels in csrc/rrt_kernels.cu.  There is no CPU path:
tensors that are not on a CUDA device raise.

Replaces (reference paths): Scene.build + theano.function (scene.py:18-52,
optimize_brightness.py:43-46) and T.grad over that graph (optimize.py:25,73).
"""
import ctypes as C
from dataclasses import dataclass, replace

import torch

from . import _native as nat


@dataclass(frozen=True)
class RenderConfig:
    """Static (non-differentiable) part of a scene: what Scene.build reads besides
    the parameter tensors (scene.py:18-52)."""
    n: int
    samples: int = 4                    # antialias_samples, scene.py:18
    shader: int = nat.SHADER_PHONG
    transpose: int = 1                  # 1 = root camera variant, 0 = orbit variant
    max_depth: float = 1.0              # DepthMapShader.maxDepth, shader.py:11
    camera_grad: int = 0
    seed: int = 0                       # in-kernel jitter seed when no jitter tensors are given
    row_begin: int = 0                  # multi-GPU row slab
    row_count: int = 0
    scene_begin: int = 0                # multi-GPU scene-batch shard: global index of scene 0 (jitter RNG key)
    cull: int = 0                       # 1: conservative per-tile object culling (bit-identical results, less work)
    no_small: int = 0                   # 1: never take the small-scene (one ray per thread) kernel (A/B, tests)
    use_records: int = 1                # 0: never precompute / TMA-stage the sweep records (A/B, tests)
    shadows: int = 0                    # 1: hard shadows (scene.py:41-45 + shape.py:85-97; include/rrt_b200.h);
                                        # 2: same, scalar pass only in the general kernel (A/B, tests)
    geom_grad_only: int = 0             # 1: RRT_FLAG_NO_MATERIAL_GRAD -- the reverse pass yields d/d w2o (+ camera) only,
                                        # material / light / look_at gradients are zero and not computed
    use_ticket: int = 1                 # 0: never fold the gradient finalisation into the render kernel (A/B, tests)
    deterministic: int = 0              # 1: RRT_FLAG_DETERMINISTIC -- gradients and loss bit-identical from run to run
                                        # (fixed-point accumulation across warps / CTAs instead of float atomics)
    pixel_threads: int = 0              # small-scene kernel thread mapping: 0 = the library's choice, 1 = force one pixel per
                                        # thread where it applies (RRT_FLAG_PIXEL_THREADS), 2 = force one ray per thread
    linear_cost: int = 0                # 1: RRT_FLAG_LINEAR_COST -- the fused entry points read `target` as a weight image W and
                                        # evaluate cost = sum(W * image) (optimize_brightness.py:51) instead of the squared error
    canonical_sweep: int = 0            # 1: RRT_FLAG_CANONICAL_SWEEP -- no conservative pre-filter in the sweep (same bits,
                                        # every pair evaluated with the reference's arithmetic; A/B, roofline accounting)

    @property
    def rows(self):
        return self.row_count if self.row_count > 0 else self.n - self.row_begin

    def slab(self, row_begin, row_count):
        return replace(self, row_begin=row_begin, row_count=row_count)


def _f32(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise nat.NativeError('%s must be a CUDA tensor (there is no CPU fallback)' % name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


RECORDS_MIN_N = 64           # from this many objects the sweep records are built once per render
                             # (rrt_build_records) and TMA-staged, instead of rebuilt in every CTA

_TICKETS = {}


def _ticket(device, num_scenes):
    """Per-(device, stream) scratch for rrt_scene.ticket (uint32 [num_scenes], zero between calls):
    lets the last CTA of each scene finalise the gradients inside the render kernel.  Calls on one
    stream are serialised, so they may share it; different streams get different scratch.  While a
    CUDA graph is being captured nothing is allocated (a memset node would be recorded): a stream
    without scratch simply takes the separate finalize launch."""
    stream = torch.cuda.current_stream(device)
    key = (str(device), stream.cuda_stream)
    t = _TICKETS.get(key)
    if t is None or t.numel() < num_scenes:
        if torch.cuda.is_current_stream_capturing():
            return None
        with torch.cuda.device(device):
            t = _TICKETS[key] = torch.zeros(max(int(num_scenes), 4096), dtype=torch.int32, device=device)
    return t


_BASE_RAYS = {}
BASE_RAYS_MAX_N = 1024       # table = 12*n*n bytes; above this the kernels evaluate the grid themselves


def base_rays(n, device):
    """Cached [n,n,3] table of Camera.make_rays' grid (scene.py:66-72) for small images,
    filled once per (n, device) by rrt_primary_rays."""
    key = (int(n), str(device))
    t = _BASE_RAYS.get(key)
    if t is None:
        with torch.cuda.device(device):
            t = torch.empty((n, n, 3), dtype=torch.float32, device=device)
            rc = nat.lib().rrt_primary_rays(int(n), t.data_ptr(), C.c_void_p(torch.cuda.current_stream(device).cuda_stream))
        nat.check(rc, 'rrt_primary_rays')
        _BASE_RAYS[key] = t
    return t


class _Tables:
    """Packed device tables + the rrt_scene descriptor that points at them."""

    def __init__(self, cfg, obj_type, w2o, material, light, camera, jitter=None, reflectivity=None):
        w2o = _f32(w2o, 'w2o')
        self.batched = w2o.dim() == 3
        if not self.batched:
            w2o = w2o.unsqueeze(0)
        self.B, self.N = int(w2o.shape[0]), int(w2o.shape[1])
        if w2o.shape[2] != nat.W2O_STRIDE:
            raise ValueError('w2o must be [..., N, 12] (rows 0..2 of shape.w2o.m)')
        self.w2o = w2o
        self.material = _f32(material, 'material').reshape(-1 if self.N else 1, self.N, nat.MAT_STRIDE)
        self.light = _f32(light, 'light').reshape(-1, nat.LIGHT_STRIDE)
        self.camera = _f32(camera, 'camera').reshape(-1, nat.CAMERA_STRIDE)
        if not obj_type.is_cuda:
            raise nat.NativeError('obj_type must be a CUDA tensor')
        self.obj_type = obj_type.to(torch.int32).contiguous()
        if self.obj_type.numel() != self.N:
            raise ValueError('obj_type must have one entry per object')
        self.cfg = cfg
        self.device = w2o.device
        self.jx = self.jy = None
        if jitter is not None:
            self.jx, self.jy = _f32(jitter[0], 'jitter_x'), _f32(jitter[1], 'jitter_y')
            per = cfg.rows * cfg.n * cfg.samples
            if self.jx.numel() not in (per, per * self.B) or self.jy.numel() != self.jx.numel():
                raise ValueError('jitter must be [rows,n,S] (shared) or [B,rows,n,S], slab-local, image index space')
        for name, t in (('material', self.material), ('light', self.light), ('camera', self.camera)):
            if t.shape[0] not in (1, self.B):
                raise ValueError('%s must be shared or have one table per scene' % name)

        d = nat.RrtScene()
        d.n, d.samples, d.num_objects, d.num_scenes = cfg.n, cfg.samples, self.N, self.B
        d.shader, d.transpose = cfg.shader, cfg.transpose
        d.row_begin, d.row_count = cfg.row_begin, cfg.row_count
        d.scene_begin = cfg.scene_begin
        # mirror bounce (RRT_FLAG_MIRROR, an extension): per-object reflectivity [N] or [B,N], a constant
        self.reflectivity = None
        if reflectivity is not None:
            self.reflectivity = _f32(reflectivity, 'reflectivity').reshape(-1, self.N)
            if self.reflectivity.shape[0] not in (1, self.B):
                raise ValueError('reflectivity must be [N] or [B, N]')
            if not cfg.transpose or cfg.camera_grad:
                raise nat.NativeError('the mirror bounce supports the root camera variant only (identity camera)')
        d.flags = ((nat.FLAG_MIRROR if self.reflectivity is not None else 0) | (nat.FLAG_CULL if cfg.cull else 0) | (nat.FLAG_NO_SMALL if cfg.no_small else 0) |
                   (nat.FLAG_SHADOWS if cfg.shadows else 0) | (nat.FLAG_SCALAR_SHADOWS if cfg.shadows == 2 else 0) |
                   (nat.FLAG_NO_MATERIAL_GRAD if cfg.geom_grad_only else 0) |
                   (nat.FLAG_CANONICAL_SWEEP if cfg.canonical_sweep else 0) |
                   (nat.FLAG_DETERMINISTIC if cfg.deterministic else 0) |
                   (nat.FLAG_LINEAR_COST if cfg.linear_cost else 0) |
                   (nat.FLAG_PIXEL_THREADS if cfg.pixel_threads == 1 else 0) |
                   (nat.FLAG_RAY_THREADS if cfg.pixel_threads == 2 else 0))
        d.max_depth, d.camera_grad, d.seed = cfg.max_depth, cfg.camera_grad, cfg.seed & 0xFFFFFFFFFFFFFFFF
        d.obj_type, d.w2o, d.material = self.obj_type.data_ptr(), self.w2o.data_ptr(), self.material.data_ptr()
        d.light, d.camera = self.light.data_ptr(), self.camera.data_ptr()
        d.w2o_scene_stride = 0 if self.B == 1 else self.N * nat.W2O_STRIDE
        d.material_scene_stride = 0 if self.material.shape[0] == 1 else self.N * nat.MAT_STRIDE
        d.light_scene_stride = 0 if self.light.shape[0] == 1 else nat.LIGHT_STRIDE
        d.camera_scene_stride = 0 if self.camera.shape[0] == 1 else nat.CAMERA_STRIDE
        if self.jx is not None:
            d.jitter_x, d.jitter_y = self.jx.data_ptr(), self.jy.data_ptr()
            per = cfg.rows * cfg.n * cfg.samples
            d.jitter_scene_stride = 0 if self.jx.numel() == per else per
        self.base = None
        if cfg.n <= BASE_RAYS_MAX_N:
            self.base = base_rays(cfg.n, self.device)
            d.base_rays = self.base.data_ptr()
        if self.reflectivity is not None:
            d.reflectivity = self.reflectivity.data_ptr()
            d.reflectivity_scene_stride = 0 if self.reflectivity.shape[0] == 1 else self.N
        self.det_ws = None
        if cfg.deterministic:
            with torch.cuda.device(self.device):
                self.det_ws = torch.empty((self.B, nat.grad_size(self.N) + 1, 2), dtype=torch.int64, device=self.device)
            d.det_workspace = self.det_ws.data_ptr()
        self.desc = d
        self.ticket = _ticket(self.device, self.B) if cfg.use_ticket else None
        if self.ticket is not None:
            d.ticket = self.ticket.data_ptr()
        self.records = None
        if self.N >= RECORDS_MIN_N and cfg.use_records:
            with torch.cuda.device(self.device):
                self.records = torch.empty(nat.record_table_floats(self.B, self.N), dtype=torch.float32, device=self.device)
                rc = nat.lib().rrt_build_records(C.byref(d), self.records.data_ptr(), self.stream())
            nat.check(rc, 'rrt_build_records')
            d.obj_records = self.records.data_ptr()

    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)


def render_forward(cfg, obj_type, w2o, material, light, camera, jitter=None, want_hit=True, want_tmin=False,
                   reflectivity=None):
    """-> image [B,rows,n,3] (or [rows,n,3] when w2o is unbatched), hit_index, tmin.
    `reflectivity` [N]: one mirror bounce (RRT_FLAG_MIRROR in include/rrt_b200.h; an extension)."""
    T = _Tables(cfg, obj_type, w2o, material, light, camera, jitter, reflectivity)
    with torch.cuda.device(T.device):
        image = torch.empty((T.B, cfg.rows, cfg.n, 3), dtype=torch.float32, device=T.device)
        hit = torch.empty((T.B, cfg.samples, cfg.rows, cfg.n), dtype=torch.int32, device=T.device) if want_hit else None
        tmin = torch.empty((T.B, cfg.samples, cfg.rows, cfg.n), dtype=torch.float32, device=T.device) if want_tmin else None
        rc = nat.lib().rrt_render_forward(C.byref(T.desc), image.data_ptr(),
                                          hit.data_ptr() if want_hit else None,
                                          tmin.data_ptr() if want_tmin else None, T.stream())
    nat.check(rc, 'rrt_render_forward')
    if not T.batched:
        image = image[0]
        hit = hit[0] if hit is not None else None
        tmin = tmin[0] if tmin is not None else None
    return image, hit, tmin


def render_backward(cfg, obj_type, w2o, material, light, camera, dl_dimage, hit_index=None, jitter=None, reflectivity=None):
    """-> flat gradient [B, N*19+21] (layout: include/rrt_b200.h)."""
    T = _Tables(cfg, obj_type, w2o, material, light, camera, jitter, reflectivity)
    dl = _f32(dl_dimage, 'dl_dimage')
    if dl.numel() != T.B * cfg.rows * cfg.n * 3:
        raise ValueError('dl_dimage must be [B, rows, n, 3]')
    with torch.cuda.device(T.device):
        grad = torch.empty((T.B, nat.grad_size(T.N)), dtype=torch.float32, device=T.device)
        if hit_index is not None:
            hit_index = hit_index.to(torch.int32).contiguous()
            if hit_index.numel() != T.B * cfg.samples * cfg.rows * cfg.n:
                raise ValueError('hit_index must be [B, S, rows, n]')
        rc = nat.lib().rrt_render_backward(C.byref(T.desc), dl.data_ptr(),
                                           hit_index.data_ptr() if hit_index is not None else None,
                                           grad.data_ptr(), T.stream())
    nat.check(rc, 'rrt_render_backward')
    return grad if T.batched else grad[0]


def render_fused_mse(cfg, obj_type, w2o, material, light, camera, target, channel_weight=None, jitter=None,
                     want_image=False, want_hit=False, reflectivity=None):
    """Fused forward + sum_c w_c*sum((image-target)^2) + reverse pass, one kernel.
    -> loss float64 [B], grad float32 [B, N*19+21], image or None, hit_index or None."""
    T = _Tables(cfg, obj_type, w2o, material, light, camera, jitter, reflectivity)
    tg = _f32(target, 'target')
    if tg.numel() != T.B * cfg.rows * cfg.n * 3:
        raise ValueError('target must be [B, rows, n, 3]')
    cw = None
    if channel_weight is not None:
        cw = (C.c_float * 3)(*[float(v) for v in channel_weight])
    with torch.cuda.device(T.device):
        # loss and grad are carved out of ONE allocation, loss first: the library then zeroes both with one
        # memset node instead of two (2.3 us of a 100 us decoder batch)
        G = nat.grad_size(T.N)
        both = torch.empty((T.B * (8 + 4 * G),), dtype=torch.uint8, device=T.device)
        loss = both[:T.B * 8].view(torch.float64)
        grad = both[T.B * 8:].view(torch.float32).view(T.B, G)
        image = torch.empty((T.B, cfg.rows, cfg.n, 3), dtype=torch.float32, device=T.device) if want_image else None
        hit = torch.empty((T.B, cfg.samples, cfg.rows, cfg.n), dtype=torch.int32, device=T.device) if want_hit else None
        rc = nat.lib().rrt_render_fused_mse(C.byref(T.desc), tg.data_ptr(), cw,
                                            image.data_ptr() if want_image else None,
                                            hit.data_ptr() if want_hit else None,
                                            loss.data_ptr(), grad.data_ptr(), T.stream())
    nat.check(rc, 'rrt_render_fused_mse')
    if not T.batched:
        loss, grad = loss[0], grad[0]
        image = image[0] if image is not None else None
        hit = hit[0] if hit is not None else None
    return loss, grad, image, hit


def slab_schedule(rows, slabs=None, heights=None):
    """Row slabs of StreamedFusedMSE -> [(first_row, height), ...] covering [0, rows) in order.
    `heights=[...]`: the caller's schedule (multiples of 4 rows keep whole CTAs); `slabs=k`: k uniform slabs;
    neither: the automatic choice -- uniform slabs of >= ~44 rows, at most 32 of them, and for tall images
    (>= 2048 rows) graded heights."""
    rows = int(rows)
    if heights is not None:
        hs = [int(h) for h in heights]
        if sum(hs) != rows or any(h <= 0 for h in hs):
            raise ValueError('heights must be positive and sum to the rows of the call')
        out, r0 = [], 0
        for h in hs:
            out.append((r0, h))
            r0 += h
        return out
    auto = slabs is None
    if slabs is None:
        # measured on C5 (tools/streamed_probe.py): 4096 rows: 16 / 24 / 32 slabs -> 24.89 / 24.67 /
        # 24.62 ms against 24.28 ms resident; a 512-row slab (one of 8 GPUs): 3 / 6 / 12 / 16 slabs ->
        # 3.53 / 3.35 / 3.31 / 3.35 ms against 3.14 ms.  => slabs of >= ~44 rows, at most 32 of them
        slabs = max(1, min(32, rows // 44))
    slabs = max(1, min(int(slabs), rows))
    per = (rows + slabs - 1) // slabs
    per = (per + 3) // 4 * 4                      # whole CTAs (4 rows each) per slab
    if auto and rows >= 2048:
        # Tall images: graded heights.  The pipeline cannot start before the FIRST slab's target has arrived
        # and is not done before the LAST slab's image has left, so the slabs at both ends are short (16, 32,
        # 64, 128 rows, mirrored at the end) and the ones in between tall (~ rows / 21 >= 128: every launch
        # has a constant cost).  tools/streamed_schedule_probe.py on C5, 4096 rows: 32 uniform
        # slabs 15.80 ms, graded ends + 128-row middle 15.65, + 192..512-row middle 15.54 (resident 15.30).
        # Short images (a rank's slab at 4-8 GPUs) keep uniform slabs: 4..16-row launches cost more than
        # they save there (512 rows: 2.19 ms uniform, 2.32 graded).
        mid = max(128, (rows // 21 + 3) // 4 * 4)
        ramp = [16, 32, 64] + ([128] if mid > 128 else [])
        left = rows - 2 * sum(ramp)
        hs = list(ramp) + [mid] * (left // mid) + ([left % mid] if left % mid else []) + ramp[::-1]
        return slab_schedule(rows, heights=hs)
    return [(r0, min(per, rows - r0)) for r0 in range(0, rows, per)]


class StreamedFusedMSE:
    """Fused forward + squared-error loss + reverse pass of ONE big image whose target (and,
    optionally, rendered image) live in pinned HOST memory: the image is cut into row slabs
    and the three legs of every slab run on their own CUDA streams --

        copy-in stream :  target slab k+1   host -> device
        kernel streams :  fused kernel on slab k        (two streams, so that the tail of
                          one slab's grid overlaps the head of the next)
        copy-out stream:  image slab k-1    device -> host

    -- so a step costs max(kernel, PCIe) instead of their sum.  Slab results are the same
    bits as the whole-image launch (rays are keyed by image row, RenderConfig.slab); the
    per-slab gradient vectors and losses are summed on the caller's stream.  Buffers,
    streams and events are created once (the equivalent of the reference's compile step,
    optimize.py:29); __call__ only enqueues work.  `slabs=None` picks the slab schedule from the
    image height (uniform slabs; graded heights from 2048 rows on), `slabs=k` asks for k uniform slabs,
    `heights=[...]` for an explicit schedule."""

    def __init__(self, cfg, num_objects, device, slabs=None, want_image=True, heights=None):
        self.cfg, self.N, self.device = cfg, int(num_objects), torch.device(device)
        self.bounds = slab_schedule(cfg.rows, slabs, heights)
        rows = cfg.rows
        K = len(self.bounds)
        with torch.cuda.device(self.device):
            self.dev_target = torch.empty((rows, cfg.n, 3), dtype=torch.float32, device=self.device)
            self.dev_image = torch.empty_like(self.dev_target) if want_image else None
            self.grads = torch.empty((K, nat.grad_size(self.N)), dtype=torch.float32, device=self.device)
            self.losses = torch.empty((K,), dtype=torch.float64, device=self.device)
            self.s_in, self.s_out = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
            self.s_k = [torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)]
            self.e_in = [torch.cuda.Event() for _ in range(K)]
            self.e_k = [torch.cuda.Event() for _ in range(K)]
            self.e_fork, self.e_out = torch.cuda.Event(), torch.cuda.Event()
            # rrt_scene.ticket scratch, one per kernel stream (their kernels overlap)
            self.tickets = [torch.zeros(16, dtype=torch.int32, device=self.device) for _ in self.s_k] if cfg.use_ticket else None
            # RRT_FLAG_DETERMINISTIC: one fixed-point workspace per kernel stream, too
            self.det_ws = [torch.empty((nat.grad_size(self.N) + 1, 2), dtype=torch.int64, device=self.device)
                           for _ in self.s_k] if cfg.deterministic else None
            torch.cuda.current_stream(self.device).synchronize()
        # fused kernel per slab (the gradient finalisation is folded into it through rrt_scene.ticket),
        # + one rrt_build_records per call from RECORDS_MIN_N objects
        self.launches_per_call = (1 if cfg.use_ticket else 2) * K + (1 if self.N >= RECORDS_MIN_N and cfg.use_records else 0)

    def __call__(self, obj_type, w2o, material, light, camera, target_host, image_host=None, channel_weight=None):
        cfg, dev = self.cfg, self.device
        if target_host.is_cuda or not target_host.is_pinned():
            raise ValueError('target_host must be a pinned host tensor (use render_fused_mse for device targets)')
        if tuple(target_host.shape) != tuple(self.dev_target.shape) or target_host.dtype != torch.float32:
            raise ValueError('target_host must be float32 [rows, n, 3]')
        if image_host is not None:
            if self.dev_image is None:
                raise ValueError('constructed with want_image=False')
            if image_host.is_cuda or not image_host.is_pinned() or tuple(image_host.shape) != tuple(self.dev_image.shape):
                raise ValueError('image_host must be a pinned host float32 [rows, n, 3] tensor')
        cw = (C.c_float * 3)(*[float(v) for v in channel_weight]) if channel_weight is not None else None
        L = nat.lib()
        cur = torch.cuda.current_stream(dev)
        # Parameter tables (dtype / layout copies) and the sweep-record table are built ONCE per
        # call, on the caller's stream, BEFORE the fork event: the record table does not depend on
        # the slab, and the side streams that TMA-load it are ordered after it by e_fork.
        T = _Tables(cfg, obj_type, w2o, material, light, camera, None)
        if T.B != 1:
            raise ValueError('StreamedFusedMSE renders one scene (use render_fused_mse for scene batches)')
        self._tables = T                              # keeps the tables alive while the side streams run
        self.e_fork.record(cur)                       # parameter + record tables are ready from here on
        for s in (self.s_in, self.s_out, *self.s_k):
            s.wait_event(self.e_fork)
        for k, (r0, rc) in enumerate(self.bounds):
            with torch.cuda.stream(self.s_in):
                self.dev_target[r0:r0 + rc].copy_(target_host[r0:r0 + rc], non_blocking=True)
                self.e_in[k].record(self.s_in)
            sk = self.s_k[k & 1]
            desc = nat.RrtScene.from_buffer_copy(T.desc)          # same tables, this slab's rows
            desc.row_begin, desc.row_count = cfg.row_begin + r0, rc
            desc.ticket = self.tickets[k & 1].data_ptr() if self.tickets is not None else None
            if self.det_ws is not None:
                desc.det_workspace = self.det_ws[k & 1].data_ptr()
            sk.wait_event(self.e_in[k])
            with torch.cuda.device(dev):
                rc_ = L.rrt_render_fused_mse(C.byref(desc), self.dev_target[r0:r0 + rc].data_ptr(), cw,
                                             self.dev_image[r0:r0 + rc].data_ptr() if self.dev_image is not None else None,
                                             None, self.losses[k:k + 1].data_ptr(), self.grads[k].data_ptr(),
                                             C.c_void_p(sk.cuda_stream))
            nat.check(rc_, 'rrt_render_fused_mse')
            self.e_k[k].record(sk)
            if image_host is not None:
                self.s_out.wait_event(self.e_k[k])
                with torch.cuda.stream(self.s_out):
                    image_host[r0:r0 + rc].copy_(self.dev_image[r0:r0 + rc], non_blocking=True)
        self.e_out.record(self.s_out)
        for e in self.e_k:                            # join
            cur.wait_event(e)
        cur.wait_event(self.e_out)
        return self.losses.sum(), self.grads.sum(0)


def split_grad(flat, N):
    """flat [..., N*19+21] -> (w2o [...,N,12], material [...,N,7], light [...,6], camera [...,15])."""
    og = flat[..., :N * nat.OBJ_GRAD_STRIDE].reshape(*flat.shape[:-1], N, nat.OBJ_GRAD_STRIDE)
    gg = flat[..., N * nat.OBJ_GRAD_STRIDE:]
    return og[..., :12], og[..., 12:19], gg[..., 0:6], gg[..., 6:21]


class _RenderFn(torch.autograd.Function):
    """image = render(w2o, material, light, camera); backward = the reverse-pass kernel."""

    @staticmethod
    def forward(ctx, w2o, material, light, camera, cfg, obj_type, jitter, reflectivity=None):
        image, hit, _ = render_forward(cfg, obj_type, w2o, material, light, camera, jitter, want_hit=True, reflectivity=reflectivity)
        ctx.cfg, ctx.obj_type, ctx.jitter, ctx.reflectivity = cfg, obj_type, jitter, reflectivity
        ctx.save_for_backward(w2o, material, light, camera, hit)
        ctx.shapes = (w2o.shape, material.shape, light.shape, camera.shape)
        return image

    @staticmethod
    def backward(ctx, dl_dimage):
        w2o, material, light, camera, hit = ctx.saved_tensors
        N = w2o.shape[-2]
        flat = render_backward(ctx.cfg, ctx.obj_type, w2o, material, light, camera, dl_dimage, hit, ctx.jitter, ctx.reflectivity)
        gw, gm, gl, gc = split_grad(flat, N)
        shp = ctx.shapes

        def fit(g, shape):
            # tables shared across a batch receive the sum over scenes
            while g.dim() > len(shape):
                g = g.sum(0)
            if g.shape != shape and g.dim() == len(shape) and shape[0] == 1 and g.shape[0] != 1:
                g = g.sum(0, keepdim=True)
            return g.reshape(shape)
        return fit(gw, shp[0]), fit(gm, shp[1]), fit(gl, shp[2]), fit(gc, shp[3]), None, None, None, None


def render(cfg, obj_type, w2o, material, light, camera, jitter=None, reflectivity=None):
    """Differentiable render: image [B,rows,n,3] / [rows,n,3] with autograd to the
    four parameter tables."""
    return _RenderFn.apply(w2o, material, light, camera, cfg, obj_type, jitter, reflectivity)


def w2o_translate_scale(centres, scales):
    """Batched, differentiable w2o rows of (translate(c) * scale(s)).inverse()
    = scale(1/s) * translate(-c)  (transform.py:35-38, 60-93): [..., 3], [..., 3] -> [..., 12].
    Off-diagonals are exact zeros (diagonal fast path of the kernels)."""
    inv = 1.0 / scales
    z = torch.zeros_like(inv[..., 0])
    b = -centres * inv
    rows = [inv[..., 0], z, z, b[..., 0], z, inv[..., 1], z, b[..., 1], z, z, inv[..., 2], b[..., 2]]
    return torch.stack(rows, dim=-1)


def render_fused_mse_loss(cfg, obj_type, w2o, material, light, camera, target, channel_weight=None, jitter=None):
    """Differentiable scalar (or [B]) squared-error loss through the fused single-kernel
    path: forward + loss + reverse pass run once; autograd just scales the stored gradient."""
    return _FusedLossFn.apply(w2o, material, light, camera, cfg, obj_type, jitter, target, channel_weight)


class _FusedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w2o, material, light, camera, cfg, obj_type, jitter, target, channel_weight):
        loss, grad, _, _ = render_fused_mse(cfg, obj_type, w2o, material, light, camera, target, channel_weight, jitter)
        ctx.save_for_backward(grad)
        ctx.N = w2o.shape[-2]
        ctx.shapes = (w2o.shape, material.shape, light.shape, camera.shape)
        return loss.float()

    @staticmethod
    def backward(ctx, g_loss):
        (grad,) = ctx.saved_tensors
        g = grad * (g_loss.unsqueeze(-1) if g_loss.dim() > 0 else g_loss)
        gw, gm, gl, gc = split_grad(g, ctx.N)
        shp = ctx.shapes

        def fit(t, shape):
            while t.dim() > len(shape):
                t = t.sum(0)
            if t.shape != shape and t.dim() == len(shape) and shape[0] == 1 and t.shape[0] != 1:
                t = t.sum(0, keepdim=True)
            return t.reshape(shape)
        return fit(gw, shp[0]), fit(gm, shp[1]), fit(gl, shp[2]), fit(gc, shp[3]), None, None, None, None, None


def measure_fp32_peak(mode=1, iters=4096):
    """FP32 pipe micro-benchmark (TFLOP/s) from the separate measurement library
    (include/rrt_b200_bench.h): mode 0 scalar FFMA, 1 packed FFMA2, 2..5 instruction mixes."""
    tf, ms = C.c_double(0), C.c_double(0)
    rc = nat.bench_lib().rrt_bench_fp32_peak(mode, iters, C.byref(tf), C.byref(ms),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise nat.NativeError('rrt_bench_fp32_peak failed (%d)' % rc)
    return tf.value, ms.value
