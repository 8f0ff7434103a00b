"""oracle_numpy -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Dense, reference-structured NumPy restatement of the hot path of
lebek/reversible-raytracer: whole-image array passes per anti-alias sample and
per shape, exactly the execution structure the reference's Theano graph has.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import this module.

The reference itself (Python 2 + Theano) cannot be imported in this image, so
this is a *restatement*; every function cites the reference file:line it follows
(paths relative to /root/reference).  It is pinned by the reference's own
artefacts in tests/test_oracle_golden.py:
  * test/test_transform.py:9-41      rotate / composition / apply known answers
  * orbit_experiments/orbit_dataset.npz[99] (+ orbit_target.npz)  forward, orbit variant
  * output/0.jpg                     forward, root variant (match_mirror.py first frame)
Gradients: PARITY UNPINNED by the reference (it ships no gradient values); the
gradient oracle is oracle_grad.py (float64 autograd over this same structure).

dtype policy: geometry (rays, matrices, det, t, masks) is
float32 like floatX=float32; shading is evaluated in float64 (Theano promotes
tuple constants such as light direction / colour to float64).

Scene description ("spec") shared by all oracles and by the tests -- plain dict:
  n            image side (reference only works for x_dims == y_dims)
  samples      anti-alias samples S                      scene.py:18
  obj_type     int32[N]   0 = Sphere, 1 = Square         shape.py:16,72
  w2o          float32[N,4,4]  shape.w2o.m               shape.py:74-75
  material     float32[N,7]  (ka, kd, ks, shininess, r, g, b)   scene.py:89-101
  light_dir    float32[3], light_int float32[3]          scene.py:78-86
  cam_o2w      float32[4,4] or None (None = root variant, scene.py:55-75;
               matrix = orbit variant, orbit_experiments/scene.py:55-80)
  look_at      float32[3]                                scene.py:57
  shader       'phong' | 'phong_nospec' | 'depth'        shader.py:9-53, orbit shader.py:45,48
  max_depth    float                                     shader.py:11
  jitter_x/y   float32[n,n,S] as drawn by build()        scene.py:24-25
"""
import numpy as np

F32 = np.float32
SPHERE, SQUARE = 0, 1


# ----------------------------------------------------------------------------
# transform.py
# ----------------------------------------------------------------------------
def identity():
    """transform.py:56-58"""
    return np.eye(4, dtype=F32), np.eye(4, dtype=F32)


def translate(x):
    """transform.py:60-75 -> (m, mInv)"""
    x = np.asarray(x, dtype=F32)
    m = np.eye(4, dtype=F32)
    m[:3, 3] = x
    mi = np.eye(4, dtype=F32)
    mi[:3, 3] = -x
    return m, mi


def scale(x):
    """transform.py:78-93 -> (m, mInv); inverse is 1/x, no numeric inversion"""
    x = np.asarray(x, dtype=F32)
    m = np.eye(4, dtype=F32)
    mi = np.eye(4, dtype=F32)
    for i in range(3):
        m[i, i] = x[i]
        mi[i, i] = F32(1.0) / x[i]
    return m, mi


def rotate(angle, axis):
    """transform.py:95-122 -> (m, mInv); angle in DEGREES, axis assumed unit,
    inverse = transpose."""
    a = np.asarray(axis, dtype=np.float64)
    radians = float(angle) * np.pi / 180.0
    s, c = np.sin(radians), np.cos(radians)
    m = np.zeros((4, 4), dtype=np.float64)
    m[0, 0] = a[0] * a[0] + (1. - a[0] * a[0]) * c
    m[0, 1] = a[0] * a[1] * (1. - c) - a[2] * s
    m[0, 2] = a[0] * a[2] * (1. - c) + a[1] * s
    m[1, 0] = a[0] * a[1] * (1. - c) + a[2] * s
    m[1, 1] = a[1] * a[1] + (1. - a[1] * a[1]) * c
    m[1, 2] = a[1] * a[2] * (1. - c) - a[0] * s
    m[2, 0] = a[0] * a[2] * (1. - c) - a[1] * s
    m[2, 1] = a[1] * a[2] * (1. - c) + a[0] * s
    m[2, 2] = a[2] * a[2] + (1. - a[2] * a[2]) * c
    m[3, 3] = 1
    m = m.astype(F32)
    return m, m.T.copy()


def compose(A, B):
    """Transform.__mul__, transform.py:35-38:  m = A.m.B.m ; mInv = B.mInv.A.mInv"""
    return (A[0] @ B[0]).astype(F32), (B[1] @ A[1]).astype(F32)


def inverse(A):
    """Transform.inverse, transform.py:32-33"""
    return A[1], A[0]


def apply_rayfield(m, origin, rays):
    """Transform.__call__(RayField), transform.py:40-47.

    NOTE the `.T` on the tensordot result reverses ALL axes, so the returned
    field is spatially transposed: rays'[a,b] = m[:3,:3] @ rays[b,a].
    """
    m = np.asarray(m, dtype=F32)
    o = np.asarray(origin, dtype=F32)
    origin2 = (m @ np.array([o[0], o[1], o[2], 1], dtype=F32))[:3]
    r4 = np.concatenate([rays, np.zeros_like(rays)[:, :, :1]], axis=2)
    r2 = np.tensordot(m, r4, [1, 2]).T[:, :, :3]
    return origin2.astype(F32), np.ascontiguousarray(r2, dtype=F32)


# ----------------------------------------------------------------------------
# scene.py: camera
# ----------------------------------------------------------------------------
def make_rays(x_dims, y_dims, sampleDist_x=None, sampleDist_y=None):
    """Camera.make_rays, scene.py:61-75 (float64 grid, normalise, cast to
    float32, THEN add the jitter in float32 without re-normalising)."""
    rays = np.dstack(np.meshgrid(np.linspace(0.5, -0.5, y_dims),
                                 np.linspace(-0.5, 0.5, x_dims), indexing='ij'))
    rays = np.dstack([rays, np.ones([y_dims, x_dims], dtype='float32')])
    rays = np.divide(rays, np.linalg.norm(rays, axis=2).reshape(
        y_dims, x_dims, 1).repeat(3, 2))
    rays = np.asarray(rays, dtype=F32)
    if sampleDist_x is not None:
        rays[:, :, 0] = rays[:, :, 0] + sampleDist_x / F32(x_dims)
    if sampleDist_y is not None:
        rays[:, :, 1] = rays[:, :, 1] + sampleDist_y / F32(y_dims)
    return np.zeros(3, dtype=F32), rays


# ----------------------------------------------------------------------------
# shape.py
# ----------------------------------------------------------------------------
def sphere_det(rays, origin):
    """Sphere._hit, shape.py:78-83"""
    pnorm = np.dot(origin, origin)
    vnorm = np.sum(rays * rays, axis=2)
    pdotv = np.tensordot(rays, origin, 1)
    return np.square(pdotv) - vnorm * (pnorm - F32(1))


def sphere_distance(w2o, origin, rays):
    """Sphere.distance, shape.py:109-126.  No t>0 test (spheres behind the
    camera still hit)."""
    o, r = apply_rayfield(w2o, origin, rays)
    with np.errstate(all='ignore'):
        pdotv = np.tensordot(r, o, 1)
        vnorm = np.sum(r * r, axis=2)
        det = sphere_det(r, o)
        d1 = (-pdotv - np.sqrt(det)) / vnorm
        d2 = (-pdotv + np.sqrt(det)) / vnorm
        dist = np.minimum(d1, d2)
        bad = (det <= 0) | np.isnan(det)
        return np.where(bad, F32(np.inf), dist).astype(F32), o, r


def sphere_normals(w2o, origin, rays):
    """Sphere.normals, shape.py:128-138: OBJECT-space normal, never mapped back."""
    dist, o, r = sphere_distance(w2o, origin, rays)
    with np.errstate(all='ignore'):
        d0 = np.where(np.isinf(dist), F32(0), dist)
        proj = o + d0[:, :, None] * r
        return proj / np.sqrt(np.sum(proj ** 2, 2))[:, :, None]


def sphere_shadow(points, light_dir):
    """Sphere.shadow, shape.py:85-97: -1 where the point is not in this (unit) sphere's
    shadow, else the first root (>= 0 means in shadow at the call site, scene.py:44-45).
    `points` = "vector from points to our center", i.e. in THIS sphere's object space."""
    y = points
    with np.errstate(all='ignore'):
        x = np.tensordot(y, -1 * normed_dir(light_dir), 1)
        decider = np.square(x) - np.sum(np.multiply(y, y), 2) + 1
        bad = np.isnan(decider) | (decider <= 0)
        return np.where(bad, -1, -x - np.sqrt(decider))


def surface_pts(w2o, origin, rays, distance):
    """Sphere.surface_pts, shape.py:100-106 with its undefined name `rays` read as
    rf.rays (the only reading that type-checks): rf.origin + stabilized * rf.rays.
    `w2o` is the frame the points are wanted in (see render(): the shadow CASTER's)."""
    o, r = apply_rayfield(w2o, origin, rays)
    stabilized = np.where(np.isinf(distance), F32(1000), distance)
    return o + stabilized[:, :, None] * r


def square_hit(rays, origin):
    """Square._hit, shape.py:25-40 (strict inequalities)."""
    with np.errstate(all='ignore'):
        not_par = rays[:, :, 2] != 0
        ts = -origin[2] / rays[:, :, 2]
        pos_t = ts > 0
        inter = origin + ts[:, :, None] * rays
        mx = (inter[:, :, 0] > -0.5) & (inter[:, :, 0] < 0.5)
        my = (inter[:, :, 1] > -0.5) & (inter[:, :, 1] < 0.5)
        mask = mx & my & pos_t & not_par
        ts = np.where(mask, ts, F32(np.inf))
    return mask, ts.astype(F32)


def square_distance(w2o, origin, rays):
    """Square.distance, shape.py:43-50"""
    o, r = apply_rayfield(w2o, origin, rays)
    mask, ts = square_hit(r, o)
    return ts, o, r


def square_normals(w2o, origin, rays):
    """Square.normals, shape.py:52-69: (0,0,+-1) by the sign of o'_z, times mask."""
    o, r = apply_rayfield(w2o, origin, rays)
    mask, ts = square_hit(r, o)
    sgn = F32(1.0) if o[2] > 0 else F32(-1.0)
    nrm = np.zeros(r.shape, dtype=F32)
    nrm[:, :, 2] = sgn
    return nrm * mask[:, :, None]


# ----------------------------------------------------------------------------
# shader.py
# ----------------------------------------------------------------------------
def normed_dir(direction):
    """Light.normed_dir, scene.py:83-86"""
    d = np.asarray(direction, dtype=np.float64)
    return d / np.sqrt(d[0] ** 2 + d[1] ** 2 + d[2] ** 2)


def phong_shade(normals, dist, material, light_dir, light_int, look_at, specular=True):
    """PhongShader.shade, shader.py:28-53 (orbit variant: specular dropped,
    orbit_experiments/shader.py:45,48).  float64 shading."""
    ka, kd, ks, sh = [np.float64(v) for v in material[:4]]
    color = np.asarray(material[4:7], dtype=np.float64)
    Lh = normed_dir(light_dir)
    n = normals.astype(np.float64)
    with np.errstate(all='ignore'):
        ndl = np.tensordot(n, -Lh, 1)
        diffuse = kd * ndl
        ph = ka + diffuse
        if specular:
            rm = 2.0 * ndl[:, :, None] * n + Lh
            rv = np.tensordot(rm, np.asarray(look_at, dtype=np.float64), 1)
            ph = ka + diffuse + ks * (rv ** sh)
        colorized = ph[:, :, None] * color[None, None, :] * \
            np.asarray(light_int, dtype=np.float64)[None, None, :]
        clipped = np.clip(colorized, 0, 1)
    return np.where(np.isinf(dist)[:, :, None], 0.0, clipped)


def depth_shade(dist, max_depth):
    """DepthMapShader.shade, shader.py:14-20 (no clip; -inf on misses)."""
    with np.errstate(all='ignore'):
        scaled = (dist.astype(np.float64) - 0) / (float(max_depth) - 0)
        return (1 - scaled)[:, :, None] * np.ones(3)


# ----------------------------------------------------------------------------
# mirror bounce (RRT_FLAG_MIRROR in include/rrt_b200.h) -- an EXTENSION, dense restatement
# ----------------------------------------------------------------------------
def mirror_shade(spec, k, origin, rays, dist, shader):
    """What the reflected rays of shape k's hits see, dense over the image like everything
    else here: -> (rgb2 float64[n,n,3], hit2 int32[n,n]).  The reference has no secondary ray
    (match_mirror.py:40,45; the hook would be scene.py:41-45), so PARITY IS UNPINNED by it; the
    semantics are those stated in include/rrt_b200.h: world normal n_w = A^T n_o / |.|,
    r = d - 2 (d.n_w) n_w from P = origin + t d, every OTHER shape in list order, nearest hit with
    t2 > 0 (strict '<'), shaded by the scene's shader as seen along r."""
    n = rays.shape[0]
    w2o = np.asarray(spec['w2o'][k], dtype=F32)
    A = w2o[:3, :3]
    o, r = apply_rayfield(w2o, origin, rays)                 # object space, IMAGE index space
    d_w = np.ascontiguousarray(rays.transpose(1, 0, 2)) if spec.get('cam_o2w') is None else rays
    with np.errstate(all='ignore'):
        t = np.where(np.isinf(dist), F32(0), dist).astype(F32)
        if spec['obj_type'][k] == SPHERE:
            p = o + t[:, :, None] * r
            n_o = p / np.sqrt(np.sum(p * p, 2))[:, :, None]
        else:
            n_o = np.zeros_like(r)
            n_o[:, :, 2] = F32(1.0) if o[2] > 0 else F32(-1.0)
        m = np.tensordot(n_o, A, 1)                          # A^T n_o per pixel
        n_w = m / np.sqrt(np.sum(m * m, 2))[:, :, None]
        dn = np.sum(d_w * n_w, 2)
        refl = (d_w - F32(2) * dn[:, :, None] * n_w).astype(F32)
        P = (np.asarray(origin, dtype=F32)[None, None, :] + t[:, :, None] * d_w).astype(F32)
        rgb2 = np.zeros((n, n, 3), dtype=np.float64)
        hit2 = np.full((n, n), -1, dtype=np.int32)
        tmin = np.full((n, n), np.inf, dtype=F32)
        for j in range(len(spec['obj_type'])):
            if j == k:
                continue
            wj = np.asarray(spec['w2o'][j], dtype=F32)
            Aj, bj = wj[:3, :3], wj[:3, 3]
            o2 = (np.tensordot(P, Aj.T, 1) + bj[None, None, :]).astype(F32)
            d2 = np.tensordot(refl, Aj.T, 1).astype(F32)
            if spec['obj_type'][j] == SPHERE:
                pd = np.sum(d2 * o2, 2)
                vn = np.sum(d2 * d2, 2)
                det = np.square(pd) - vn * (np.sum(o2 * o2, 2) - F32(1))
                t2 = (-pd - np.sqrt(det)) / vn
                t2 = np.where((det <= 0) | np.isnan(det), F32(np.inf), t2).astype(F32)
                p2 = o2 + np.where(np.isinf(t2), F32(0), t2)[:, :, None] * d2
                nrm2 = p2 / np.sqrt(np.sum(p2 * p2, 2))[:, :, None]
            else:
                t2 = -o2[:, :, 2] / d2[:, :, 2]
                inter = o2 + t2[:, :, None] * d2
                mask = (inter[:, :, 0] > -0.5) & (inter[:, :, 0] < 0.5) & (inter[:, :, 1] > -0.5) & (inter[:, :, 1] < 0.5) & \
                    (t2 > 0) & (d2[:, :, 2] != 0)
                t2 = np.where(mask, t2, F32(np.inf)).astype(F32)
                nrm2 = np.zeros_like(d2)
                nrm2[:, :, 2] = np.where(o2[:, :, 2] > 0, F32(1), F32(-1))
            t2 = np.where(t2 > 0, t2, F32(np.inf)).astype(F32)          # secondary hits must lie ahead
            shad2 = phong_shade(nrm2, t2, spec['material'][j], spec['light_dir'], spec['light_int'],
                                spec['look_at'], specular=(shader == 'phong'))
            take = (t2 < tmin) & ~np.isinf(dist)
            rgb2 = np.where(take[:, :, None], shad2, rgb2)
            tmin = np.where(take, t2, tmin)
            hit2[take] = j
    return rgb2, hit2


# ----------------------------------------------------------------------------
# scene.py: Scene.build
# ----------------------------------------------------------------------------
def draw_jitter(n, samples, rng):
    """scene.py:24-25: x array drawn first, then y, from one RandomState."""
    jx = np.asarray(rng.random_sample((n, n, samples)), dtype=F32)
    jy = np.asarray(rng.random_sample((n, n, samples)), dtype=F32)
    return jx, jy


def render(spec, return_aux=True):
    """Scene.build, scene.py:18-52 evaluated eagerly.

    Returns image float64[n,n,3] and, if return_aux, hit_index int32[S,n,n]
    (-1 = background, else index into the shape list) and tmin float32[S,n,n].
    """
    n, S = int(spec['n']), int(spec['samples'])
    jx, jy = spec['jitter_x'], spec['jitter_y']
    image = np.zeros((n, n, 3), dtype=np.float64)
    hit_index = np.full((S, n, n), -1, dtype=np.int32)
    tmins = np.full((S, n, n), np.inf, dtype=F32)
    shader = spec['shader']
    shadows = bool(spec.get('shadows', 0))
    shadow_mask = np.zeros((S, n, n), dtype=bool)
    hit2_index = np.full((S, n, n), -1, dtype=np.int32)       # mirror bounce: what the reflected ray hit
    for s in range(S):
        sdx = (jx[:, :, s] + F32(s)) / F32(S)          # scene.py:31
        sdy = (jy[:, :, s] + F32(s)) / F32(S)          # scene.py:32
        origin, rays = make_rays(n, n, sdx, sdy)
        if spec.get('cam_o2w') is not None:            # orbit_experiments/scene.py:80
            origin, rays = apply_rayfield(spec['cam_o2w'], origin, rays)
        img_s = np.zeros((n, n, 3), dtype=np.float64)
        min_d = np.full((n, n), np.inf, dtype=F32)
        for k in range(len(spec['obj_type'])):
            w2o = spec['w2o'][k]
            if spec['obj_type'][k] == SPHERE:
                dist, _, _ = sphere_distance(w2o, origin, rays)
                nrm = sphere_normals(w2o, origin, rays) if shader != 'depth' else None
            else:
                dist, _, _ = square_distance(w2o, origin, rays)
                nrm = square_normals(w2o, origin, rays) if shader != 'depth' else None
            if shader == 'depth':
                shad = depth_shade(dist, spec['max_depth'])
            else:
                shad = phong_shade(nrm, dist, spec['material'][k], spec['light_dir'],
                                   spec['light_int'], spec['look_at'],
                                   specular=(shader == 'phong'))
            hit2_k = None
            if spec.get('reflectivity') is not None:                 # one mirror bounce (extension)
                kr = float(spec['reflectivity'][k])
                rgb2, hit2_k = mirror_shade(spec, k, origin, rays, dist, shader)
                shad = np.where(np.isinf(dist)[:, :, None], shad, (1.0 - kr) * shad + kr * rgb2)
            in_shadow = np.zeros((n, n), dtype=bool)
            if shadows:
                # scene.py:41-45 (commented out in the reference): for each shape != obj draw
                # its shadow on obj.  Points are taken in the CASTER's frame -- the reading
                # under which Sphere.shadow's "vector from points to our center" holds;
                # Square has no shadow method, so squares cast none.
                for k2 in range(len(spec['obj_type'])):
                    if k2 == k or spec['obj_type'][k2] != SPHERE:
                        continue
                    pts = surface_pts(spec['w2o'][k2], origin, rays, dist)
                    lit = sphere_shadow(pts, spec['light_dir']) < 0
                    shad = np.where(lit[:, :, None], shad, 0.0)
                    in_shadow |= ~lit
            take = dist < min_d                                      # scene.py:46 (strict)
            img_s = np.where(take[:, :, None], shad, img_s)
            min_d = np.where(take, dist, min_d)                      # scene.py:47
            hit_index[s][take] = k
            shadow_mask[s][take] = in_shadow[take]
            if hit2_k is not None:
                hit2_index[s][take] = hit2_k[take]
        tmins[s] = min_d
        image = image + img_s                                        # scene.py:49
    image = image / S                                                # scene.py:50
    if return_aux:
        if shadows:
            hit_index = np.where(shadow_mask & (hit_index >= 0), hit_index | 0x40000000, hit_index)
        if spec.get('reflectivity') is not None and return_aux == 'mirror':
            return image, hit_index, tmins, hit2_index
        return image, hit_index, tmins
    return image
