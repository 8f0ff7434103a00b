"""Transform algebra on torch tensors -- host-side mirror of the reference's
transform.py (same names, argument meaning and composition rules):

    Transform(m, mInv), .inverse(), A * B, A(RayField)      transform.py:27-47
    identity / translate / scale / rotate                     transform.py:56-122
    RayField(origin, directions)                              transform.py:21-24

The matrices are the DIFFERENTIABLE link from user parameters (torch tensors with
requires_grad) to the `w2o` table the kernels consume: torch autograd chains
d/d w2o (produced by the reverse-pass kernel) back into translate/scale/rotate
arguments, exactly where Theano's T.grad did in the reference.  No numeric matrix
inversion anywhere: every primitive carries its analytic inverse.
"""
import math

import numpy as np
import torch

_DEFAULT_DEVICE = None


def default_device():
    """Device used for constants given as tuples/lists/NumPy arrays."""
    global _DEFAULT_DEVICE
    if _DEFAULT_DEVICE is None:
        _DEFAULT_DEVICE = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() \
            else torch.device('cpu')
    return _DEFAULT_DEVICE


def set_default_device(device):
    global _DEFAULT_DEVICE
    _DEFAULT_DEVICE = torch.device(device)


_CONST_CACHE = {}
_CONST_CACHE_MAX = 4096


def as_tensor(x, device=None):
    """T.as_tensor_variable equivalent: tensors pass through (keeping autograd; non-float32
    tensors are cast), everything else becomes a float32 constant on the default device.
    Constants are cached BY VALUE: a loss closure that rebuilds its scene on every call, the way
    the reference's decoders do (orbit_experiments/test_optimization.py:17-44:
    `translate(center2) * scale((6, 6, 6))`, `Material((0.9, 0, 0), ...)` inside `scene()`),
    uploads each constant once and can then be captured into a CUDA graph (no host->device
    copy on the steady-state path).  The cached tensors are shared: never modify them in place."""
    if isinstance(x, torch.Tensor):
        return x if x.dtype == torch.float32 else x.float()
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
    dev = torch.device(device) if device is not None else default_device()
    if a.size > 64:                       # images / targets: not worth hashing, not constants of the algebra
        return torch.as_tensor(a, device=dev)
    key = (a.tobytes(), a.shape, str(dev))
    t = _CONST_CACHE.get(key)
    if t is None:
        if len(_CONST_CACHE) >= _CONST_CACHE_MAX:
            _CONST_CACHE.clear()
        t = _CONST_CACHE[key] = torch.as_tensor(a, device=dev)
    return t


class _Arg(object):
    """One argument of a primitive transform: a LIVE parameter (any torch tensor handed in by the
    user -- re-read on every evaluation, cast to float32 lazily so that in-place optimiser updates of
    a float64 / half parameter are seen, like a theano.shared variable) or a constant (host value
    kept for the chain compiler, device tensor from the by-value cache).

    A live parameter that is a VIEW OF A LEAF tensor -- the reference's idiom
    `translate(obj_param[:3]) * scale(obj_param[3:])` (test_balls.py:27, autoencoder.py:60) -- is kept
    as (leaf, size, stride, offset) and the view is taken AGAIN on every evaluation.  Like indexing a
    theano.shared symbolically, the view then follows the leaf whatever happened in between: it tracks
    gradients even if `requires_grad_` was set after the slice was written, and its autograd nodes are
    created inside the optimiser step (on the step's stream) instead of being pinned to the stream of
    the construction site, which is what lets such a closure be captured into a CUDA graph."""
    __slots__ = ('_src', '_view', 'host', 'param', 'key')

    def __init__(self, x, device=None):
        self.param = isinstance(x, torch.Tensor)
        self._view = None
        if self.param:
            self.host = None
            base = x._base if x._is_view() else None
            if base is not None and base.is_leaf and base.layout == torch.strided and x.layout == torch.strided:
                self._view = (base, tuple(x.shape), tuple(x.stride()), int(x.storage_offset()))
                self._src = None
                self.key = ('view', id(base)) + self._view[1:]      # parameter identity for the chain compiler
            else:
                self._src = x
                self.key = ('tensor', id(x))
        else:
            self.host = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
            self._src = as_tensor(self.host, device=device)
            self.key = None

    @property
    def src(self):
        if self._view is not None:
            base, size, stride, offset = self._view
            return base.as_strided(size, stride, offset)
        return self._src

    @property
    def t(self):
        """float32 view of the current value"""
        x = self.src
        return x if x.dtype == torch.float32 else x.float()

    @property
    def device(self):
        return self._view[0].device if self._view is not None else self._src.device


class Point(object):
    """transform.py:6-8"""

    def __init__(self, p):
        self.p = p


class PointField(object):
    """transform.py:11-13"""

    def __init__(self, pf):
        self.pf = pf


class VectorField(object):
    """transform.py:16-18"""

    def __init__(self, vf):
        self.vf = vf


class RayField(object):
    """transform.py:21-24"""

    def __init__(self, origin, directions):
        self.rays = as_tensor(directions)
        self.origin = as_tensor(origin, device=self.rays.device)


class Transform(object):
    """transform.py:27-47.

    The reference's matrices are SYMBOLIC: they are re-evaluated from the current
    values of the theano.shared parameters on every call of the compiled function.
    To keep that behaviour in eager PyTorch a Transform built from torch tensors is
    LAZY: `.m` / `.mInv` are recomputed from the live parameter tensors on every
    access (so in-place optimiser updates are seen, and every render gets a fresh
    autograd graph).  Transforms built from constants (tuples, lists, NumPy) are
    evaluated once and cached.
    """

    def __init__(self, m=None, mInv=None, _fm=None, _fmInv=None, _dynamic=False, _expr=None):
        self._m, self._mInv = m, mInv
        self._fm, self._fmInv = _fm, _fmInv
        self._dynamic = _dynamic
        # structure of the expression, for the native parameter->matrix chain (chain.py):
        #   ('T'|'S', _Arg) | ('R', _Arg angle, _Arg axis) | ('mul', A, B) | ('inv', A)
        #   | ('I',) | None (explicit matrices: not expressible, torch path only)
        self._expr = _expr

    @property
    def m(self):
        if self._fm is None:
            return self._m
        if self._dynamic:
            return self._fm()
        if self._m is None:
            self._m = self._fm()
        return self._m

    @property
    def mInv(self):
        if self._fmInv is None:
            return self._mInv
        if self._dynamic:
            return self._fmInv()
        if self._mInv is None:
            self._mInv = self._fmInv()
        return self._mInv

    def inverse(self):
        return Transform(self._mInv, self._m, self._fmInv, self._fm, self._dynamic,
                         _expr=None if self._expr is None else ('inv', self))

    def __mul__(self, other):
        # transform.py:35-38: m = A.m . B.m ; mInv = B.mInv . A.mInv.  Explicit
        # multiply-sum instead of torch.matmul: never TF32, and exact zeros stay exact
        # (the kernels' diagonal fast path keys on them).
        a, b = self, other
        return Transform(_fm=lambda: _mm4(a.m, b.m.to(a.m.device)),
                         _fmInv=lambda: _mm4(b.mInv, a.mInv.to(b.mInv.device)),
                         _dynamic=a._dynamic or b._dynamic,
                         _expr=None if (a._expr is None or b._expr is None) else ('mul', a, b))

    def __call__(self, x):
        """Apply to a RayField.  Like the reference (transform.py:44-46) the `.T` on the
        tensordot result reverses all axes, so the returned ray field is SPATIALLY
        TRANSPOSED: rays'[a,b] = m[:3,:3] @ rays[b,a].  Dense helper for API
        compatibility; the render kernels fold this into their index mapping."""
        if isinstance(x, RayField):
            o, r = x.origin, x.rays
            m = self.m
            one = torch.ones(1, dtype=o.dtype, device=o.device)
            origin = (m.to(o.device) * torch.cat([o[:3], one])[None, :]).sum(1)[:3]
            rays = torch.cat([r, torch.zeros_like(r)[:, :, :1]], dim=2)
            rays = torch.tensordot(m.to(r.device), rays, dims=([1], [2])).permute(2, 1, 0)[:, :, :3]
            return RayField(origin, rays)
        raise TypeError('Transform can only be applied to a RayField')


def _mm4(a, b):
    return (a.unsqueeze(-1) * b.unsqueeze(-3)).sum(-2)


def _eye(device):
    return torch.eye(4, dtype=torch.float32, device=device)


def identity():
    """transform.py:56-58"""
    return Transform(_eye(default_device()), _eye(default_device()), _expr=('I',))


def _place(device, entries):
    """4x4 identity with differentiable scalar entries placed at (r, c)."""
    m = _eye(device)
    rows = []
    for r in range(4):
        cols = []
        for c in range(4):
            cols.append(entries[(r, c)].reshape(()) if (r, c) in entries else m[r, c])
        rows.append(torch.stack(cols))
    return torch.stack(rows)


def translate(x):
    """transform.py:60-75"""
    a = _Arg(x)
    return Transform(_fm=lambda: _translation(a.t, 1.0), _fmInv=lambda: _translation(a.t, -1.0),
                     _dynamic=a.param, _expr=('T', a))


def _translation(x, sign):
    if sign < 0:
        x = -x
    return _place(x.device, {(0, 3): x[0], (1, 3): x[1], (2, 3): x[2]})


def scale(x):
    """transform.py:78-93 (inverse is 1/x)"""
    a = _Arg(x)
    return Transform(_fm=lambda: _scaling(a.t, False), _fmInv=lambda: _scaling(a.t, True),
                     _dynamic=a.param, _expr=('S', a))


def _scaling(x, inverse):
    if inverse:
        return _place(x.device, {(0, 0): 1. / x[0], (1, 1): 1. / x[1], (2, 2): 1. / x[2]})
    return _place(x.device, {(0, 0): x[0], (1, 1): x[1], (2, 2): x[2]})


def rotate(angle, axis):
    """transform.py:95-122: angle in DEGREES about an (assumed unit) axis; the
    inverse is the transpose."""
    ax = _Arg(axis)
    an = _Arg(angle, device=ax.device)
    return Transform(_fm=lambda: _rotation(an.t, ax.t.to(an.device)), _fmInv=lambda: _rotation(an.t, ax.t.to(an.device)).t(),
                     _dynamic=an.param or ax.param, _expr=('R', an, ax))


def _rotation(angle, a):
    radians = angle * (math.pi / 180.0)
    s, c = torch.sin(radians), torch.cos(radians)
    e = {
        (0, 0): a[0] * a[0] + (1. - a[0] * a[0]) * c,
        (0, 1): a[0] * a[1] * (1. - c) - a[2] * s,
        (0, 2): a[0] * a[2] * (1. - c) + a[1] * s,
        (1, 0): a[0] * a[1] * (1. - c) + a[2] * s,
        (1, 1): a[1] * a[1] + (1. - a[1] * a[1]) * c,
        (1, 2): a[1] * a[2] * (1. - c) - a[0] * s,
        (2, 0): a[0] * a[2] * (1. - c) - a[1] * s,
        (2, 1): a[1] * a[2] * (1. - c) + a[0] * s,
        (2, 2): a[2] * a[2] + (1. - a[2] * a[2]) * c,
    }
    return _place(a.device, e)
