"""Shapes -- host-side mirror of the reference's shape.py (Sphere, Square).

A shape is a unit sphere / unit square in OBJECT space placed by its `o2w`
Transform; it stores `w2o = o2w.inverse()` (shape.py:19-23, 73-76).  Scene.build
packs `w2o.m[:3,:]` and the material into the device tables the kernels read.

The dense per-shape methods (`distance`, `normals`, `_hit`) are kept for API
compatibility and evaluate with torch ops on whatever device the ray field lives
on; they are NOT the render path -- Scene.build() always runs the CUDA kernels.
"""
import torch

from . import _native as nat
from .transform import RayField  # noqa: F401

_INF = float('inf')


class Shape(object):
    kind = None

    def __init__(self, o2w, material):
        self.o2w = o2w
        self.w2o = o2w.inverse()
        self.material = material

    def setTransform(self, o2w):
        """shape.py:12-14"""
        self.o2w = o2w
        self.w2o = o2w.inverse()


class Square(Shape):
    """Square on the object xy-plane, vertices (+-0.5, +-0.5, 0), normal (0,0,1)
    (shape.py:16-23)."""
    kind = nat.OBJ_SQUARE

    def _hit(self, rays, origin):
        """shape.py:25-40 (strict inequalities)."""
        not_par = rays[:, :, 2] != 0
        ts = -origin[2] / rays[:, :, 2]
        inter = origin + ts.unsqueeze(-1) * rays
        mx = (inter[:, :, 0] > -0.5) & (inter[:, :, 0] < 0.5)
        my = (inter[:, :, 1] > -0.5) & (inter[:, :, 1] < 0.5)
        mask = mx & my & (ts > 0) & not_par
        ts = torch.where(mask, ts, torch.full_like(ts, _INF))
        return mask, ts

    def distance(self, rayField):
        """shape.py:43-50"""
        rf = self.w2o(rayField)
        return self._hit(rf.rays, rf.origin)[1]

    def normals(self, rayField):
        """shape.py:52-69"""
        rf = self.w2o(rayField)
        mask, _ = self._hit(rf.rays, rf.origin)
        sgn = torch.where(rf.origin[2] > 0, 1.0, -1.0)
        norm = torch.zeros_like(rf.rays)
        norm[:, :, 2] = sgn
        return norm * mask.unsqueeze(-1)


class Sphere(Shape):
    """Unit sphere in object space (shape.py:72-76)."""
    kind = nat.OBJ_SPHERE

    def _hit(self, rays, origin):
        """shape.py:78-83: the discriminant."""
        pnorm = torch.dot(origin, origin)
        vnorm = (rays * rays).sum(2)
        pdotv = (rays * origin).sum(2)
        return pdotv * pdotv - vnorm * (pnorm - 1)

    def distance(self, rayField):
        """shape.py:109-126: first root, +inf where det <= 0 or NaN; no t>0 test."""
        rf = self.w2o(rayField)
        pdotv = (rf.rays * rf.origin).sum(2)
        vnorm = (rf.rays * rf.rays).sum(2)
        det = self._hit(rf.rays, rf.origin)
        safe = torch.sqrt(torch.clamp(det, min=0))
        dist = torch.minimum((-pdotv - safe) / vnorm, (-pdotv + safe) / vnorm)
        bad = (det <= 0) | torch.isnan(det)
        return torch.where(bad, torch.full_like(dist, _INF), dist)

    def shadow(self, points, lights):
        """shape.py:85-97 (dense helper; the kernels' RRT_FLAG_SHADOWS pass follows the same formula per
        winning ray): for object-space `points` [x, y, 3] the distance along -Lhat at which the unit sphere is
        entered, -1 where the line misses it (decider <= 0 or NaN) -- "in shadow" where the result is >= 0."""
        y = points
        x = torch.tensordot(y, -1.0 * lights[0].normed_dir().to(y.device), dims=1)
        decider = x * x - (y * y).sum(2) + 1
        bad = torch.isnan(decider) | (decider <= 0)
        return torch.where(bad, torch.full_like(x, -1.0), -x - torch.sqrt(torch.clamp(decider, min=0)))

    def surface_pts(self, rayField):
        """shape.py:100-106: object-space hit points, rays that miss parked at distance 1000.  (The
        reference's body refers to an undefined name `rays`; the evident intent, rf.rays, is used.)"""
        rf = self.w2o(rayField)
        distance = self.distance(rayField)
        stabilized = torch.where(torch.isinf(distance), torch.full_like(distance, 1000.0), distance)
        return rf.origin + stabilized.unsqueeze(-1) * rf.rays

    def normals(self, rayField):
        """shape.py:128-138: object-space normal (never mapped back to world)."""
        rf = self.w2o(rayField)
        dist = self.distance(rayField)
        dist = torch.where(torch.isinf(dist), torch.zeros_like(dist), dist)
        proj = rf.origin + dist.unsqueeze(-1) * rf.rays
        return proj / torch.sqrt((proj ** 2).sum(2)).unsqueeze(-1)
