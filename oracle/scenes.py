"""oracle.scenes -- TEST INFRASTRUCTURE ONLY.

Builds the "spec" dicts (see oracle_numpy.py) of the reference's own workloads
with the ORACLE's transform algebra, so the oracle can be pinned against the
reference's golden artefacts without touching the product package.
"""
import numpy as np
from . import oracle_numpy as on

F32 = np.float32


def _mat(color, ks, kd, ka, sh):
    """Material(color, ks, kd, ka, shininess), scene.py:89-101 -> packed row."""
    return np.array([ka, kd, ks, sh, color[0], color[1], color[2]], dtype=F32)


def _chain(*ts):
    out = ts[0]
    for t in ts[1:]:
        out = on.compose(out, t)
    return out


def spec_from(n, samples, shapes, light, shader, cam=None, look_at=(0, 0, 1.),
              max_depth=1.0, seed=0):
    """shapes: list of (type, o2w(m,mInv), material row)."""
    rng = np.random.RandomState(seed)
    jx, jy = on.draw_jitter(n, samples, rng)
    return dict(
        n=n, samples=samples,
        obj_type=np.array([s[0] for s in shapes], dtype=np.int32),
        w2o=np.stack([on.inverse(s[1])[0] for s in shapes]).astype(F32),
        material=np.stack([s[2] for s in shapes]).astype(F32),
        light_dir=np.asarray(light[0], dtype=F32), light_int=np.asarray(light[1], dtype=F32),
        cam_o2w=None if cam is None else np.asarray(cam[0], dtype=F32),
        look_at=np.asarray(look_at, dtype=F32),
        shader=shader, max_depth=float(max_depth), jitter_x=jx, jitter_y=jy)


def optimize_brightness(n=128, samples=4, seed=0, c1=(-.5, -.5, 4), c2=(.5, .5, 4)):
    """C1: optimize_brightness.py:19-38"""
    m1 = _mat((0.2, 0.9, 0.4), 0.3, 0.7, 0.5, 50.)
    m2 = _mat((0.87, 0.1, 0.507), 0.3, 0.9, 0.4, 50.)
    t1 = on.translate(c1)
    t2 = _chain(on.translate(c2), on.rotate(90, (0, 0, 1)), on.scale((1, 2, 1.5)))
    return spec_from(n, samples, [(on.SPHERE, t1, m1), (on.SPHERE, t2, m2)],
                     ((-1., -1., 2.), (0.961, 1., 0.87)), 'phong', seed=seed)


def test_balls(n=32, samples=4, seed=0, p1=(0, 0, 3, .5, .5, .5), p2=(0, 0, 3, .5, .5, .5)):
    """C2: test_balls.py:22-44 (DepthMapShader(6.1); capsule bias capsule.py:10)"""
    m1 = _mat((0.2, 0.9, 0.4), 0.3, 0.7, 0.5, 50.)
    shapes = [(on.SPHERE, on.compose(on.translate(p[:3]), on.scale(p[3:])), m1) for p in (p1, p2)]
    return spec_from(n, samples, shapes, ((-1., -1., 2.), (0.961, 1., 0.87)), 'depth',
                     max_depth=6.1, seed=seed)


def match_mirror(n=128, samples=4, seed=0, c1=(-.5, -.5, 4), c2=(.5, .5, 4)):
    """C3: match_mirror.py:16-33"""
    m1 = _mat((0.2, 0.9, 0.4), 0.3, 0.7, 0.5, 50.)
    m2 = _mat((0.87, 0.1, 0.507), 0.3, 0.9, 0.4, 50.)
    sq = on.compose(on.translate((0, 0, 3)), on.rotate(50, [0., 1., 0.]))
    return spec_from(n, samples, [(on.SPHERE, on.translate(c1), m1),
                                  (on.SPHERE, on.translate(c2), m2),
                                  (on.SQUARE, sq, m2)],
                     ((-1., -1., 2.), (1., 0.87, 0.961)), 'phong', seed=seed)


def orbit(centre, view, n=64, samples=4, seed=0):
    """C4 / planet_orbit.py:21-42: two spheres, light (0,0,1), Phong without
    specular (orbit_experiments/shader.py:45,48), cameras at y=+-2.5."""
    m1 = _mat((0.0, 0.9, 0.0), 0.3, 0.7, 0.5, 50.)
    m2 = _mat((0.9, 0.0, 0.0), 0.3, 0.9, 0.4, 50.)
    shapes = [(on.SPHERE, on.compose(on.translate(centre), on.scale((4., 4., 4.))), m1),
              (on.SPHERE, on.compose(on.translate((0, 0, 48)), on.scale((6, 6, 6))), m2)]
    cam = on.translate((0, 2.5 if view == 0 else -2.5, 0))
    return spec_from(n, samples, shapes, ((0., 0., 1.), (1., 1., 1.)), 'phong_nospec',
                     cam=cam, look_at=(0, 0, 1.), seed=seed)


def stress(n=256, num_objects=64, samples=4, seed=1234, general=False, jitter_seed=4321):
    """C5 / C5g (synthetic): random spheres in the view frustum."""
    rng = np.random.RandomState(seed)
    z = rng.uniform(8, 16, num_objects)
    x = rng.uniform(-0.475, 0.475, num_objects) * z
    y = rng.uniform(-0.475, 0.475, num_objects) * z
    if general:
        sc = rng.uniform(0.10, 0.25, (num_objects, 3))
        ang = rng.uniform(0, 180, num_objects)
        ax = rng.normal(size=(num_objects, 3))
        ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    else:
        r = rng.uniform(0.10, 0.25, num_objects)
        sc = np.stack([r, r, r], 1)
    col = rng.uniform(0.1, 1.0, (num_objects, 3))
    ka = rng.uniform(.1, .5, num_objects)
    kd = rng.uniform(.5, .9, num_objects)
    shapes = []
    for k in range(num_objects):
        t = on.translate((x[k], y[k], z[k]))
        if general:
            t = on.compose(t, on.rotate(ang[k], ax[k]))
        t = on.compose(t, on.scale(sc[k]))
        shapes.append((on.SPHERE, t, _mat(col[k], 0.3, kd[k], ka[k], 50.)))
    return spec_from(n, samples, shapes, ((-1., -1., 2.), (0.961, 1., 0.87)), 'phong',
                     seed=jitter_seed)


def shadow_scene(n=64, samples=4, seed=0, shader='phong', general=False):
    """Shadows (scene.py:41-45 + shape.py:85-97, not live in the reference):
    two small spheres between the light and a big sphere / a big square, light like C1."""
    m1 = _mat((0.2, 0.9, 0.4), 0.3, 0.7, 0.5, 50.)
    m2 = _mat((0.87, 0.1, 0.507), 0.3, 0.9, 0.4, 50.)
    m3 = _mat((0.5, 0.5, 0.9), 0.2, 0.8, 0.4, 30.)
    big = _chain(on.translate((-0.3, -0.3, 5.0)), on.scale((1.2, 1.2, 1.2)))
    c1 = _chain(on.translate((0.25, 0.2, 3.6)), on.scale((0.3, 0.3, 0.3)))
    c2 = on.translate((0.9, -0.8, 4.0))
    if general:
        c2 = _chain(c2, on.rotate(30, (0, 0, 1)), on.scale((0.5, 0.25, 0.4)))
    else:
        c2 = _chain(c2, on.scale((0.35, 0.35, 0.35)))
    wall = _chain(on.translate((0, 0, 6.5)), on.scale((7, 7, 1)))
    spec = spec_from(n, samples, [(on.SPHERE, c1, m2), (on.SPHERE, big, m1), (on.SQUARE, wall, m3), (on.SPHERE, c2, m2)],
                     ((-1., -1., 2.), (0.961, 1., 0.87)), shader, max_depth=8.0, seed=seed)
    spec['shadows'] = 1
    return spec


def mirror_scene(n=64, samples=4, seed=0, general=False, reflectivity=(0.0, 0.6, 0.8, 0.3)):
    """Mirror bounce (RRT_FLAG_MIRROR; an extension -- the reference has no secondary ray): match_mirror.py's
    two spheres (the second partly reflective) in front of a reflective square 'mirror', plus a small
    sphere only visible through reflections.  Root camera variant."""
    m1 = _mat((0.2, 0.9, 0.4), 0.3, 0.7, 0.5, 50.)
    m2 = _mat((0.87, 0.1, 0.507), 0.3, 0.9, 0.4, 50.)
    m3 = _mat((0.6, 0.6, 0.9), 0.2, 0.6, 0.3, 30.)
    s1 = on.translate((-.5, -.5, 4))
    s2 = on.translate((.6, .5, 4.2))
    if general:
        s2 = _chain(s2, on.rotate(30, (0, 0, 1)), on.scale((1.1, 0.7, 0.9)))
    sq = _chain(on.translate((0.2, 0, 6.0)), on.rotate(25, [0., 1., 0.]), on.scale((6, 6, 1)))
    s3 = _chain(on.translate((-1.6, 0.9, 2.5)), on.scale((0.5, 0.5, 0.5)))
    spec = spec_from(n, samples, [(on.SPHERE, s1, m1), (on.SPHERE, s2, m2), (on.SQUARE, sq, m3), (on.SPHERE, s3, m1)],
                     ((-1., -1., 2.), (1., 0.87, 0.961)), 'phong', seed=seed)
    spec['reflectivity'] = np.asarray(reflectivity, dtype=F32)
    return spec
