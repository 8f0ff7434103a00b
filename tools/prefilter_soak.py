"""One-off soak of the conservative pre-filter: many random scenes (scales 10^-2.5..10^1.2,
rotations, anisotropy, spheres behind / around the camera, rotated cameras, n up to 160) rendered with the
pre-filter sweep and with RRT_FLAG_CANONICAL_SWEEP must agree in every bit of hit_index / tmin / image.
usage: python tools/prefilter_soak.py [first_seed] [count]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from dataclasses import replace
import numpy as np, torch
from oracle import oracle_c as oc, oracle_numpy as on, scenes
from reversible_raytracer_b200 import render as R
from helpers import to_device
first, count = (int(sys.argv[1]) if len(sys.argv) > 1 else 5000), (int(sys.argv[2]) if len(sys.argv) > 2 else 200)
dev = torch.device('cuda')
bad = 0
for seed in range(first, first + count):
    rng = np.random.RandomState(seed)
    N = int(rng.choice([64, 100, 257, 600, 1100]))
    n = int(rng.choice([24, 40, 96, 160]))
    shapes = []
    for k in range(N):
        z = rng.uniform(-4, 30)
        c = (rng.uniform(-0.6, 0.6) * abs(z) - 0.1, rng.uniform(-0.6, 0.6) * abs(z) + 0.1, z)
        t = on.translate(c)
        if rng.rand() < 0.5:
            ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
            t = on.compose(t, on.rotate(rng.uniform(0, 360), ax))
        sc = 10.0 ** rng.uniform(-2.5, 1.2, 3) if rng.rand() < 0.5 else np.full(3, 10.0 ** rng.uniform(-2.5, 1.2))
        t = on.compose(t, on.scale(sc))
        shapes.append((on.SPHERE, t, scenes._mat(rng.uniform(0.1, 1, 3), 0.3, 0.7, 0.4, 50.)))
    cam = None
    if seed % 3 == 1:
        cam = on.compose(on.translate(rng.uniform(-1, 1, 3)), on.rotate(rng.uniform(-40, 40), (0, 1, 0)))
    spec = scenes.spec_from(n, int(rng.choice([1, 2, 4, 8])), shapes, ((-1., -1., 2.), (0.961, 1., 0.87)), 'phong', cam=cam, seed=seed)
    ps = oc.PackedScene.from_spec(spec, camera_grad=0)
    cfg, ot, w2o, mat, light, cam_t, jit = to_device(ps, dev)
    cfg = replace(cfg, no_small=1, use_records=1)
    a = R.render_forward(replace(cfg, canonical_sweep=0), ot, w2o, mat, light, cam_t, jit, want_hit=True, want_tmin=True)
    b = R.render_forward(replace(cfg, canonical_sweep=1), ot, w2o, mat, light, cam_t, jit, want_hit=True, want_tmin=True)
    ok = torch.equal(a[1], b[1]) and torch.equal(a[2].view(torch.int32), b[2].view(torch.int32)) and \
        torch.equal(a[0].view(torch.int32), b[0].view(torch.int32))
    if ok and os.environ.get('ORACLE') and n <= 40:          # also against the canonical-order C oracle (CPU)
        img_o, hit_o, tmin_o = oc.render_forward(ps)
        ok = np.array_equal(a[1].cpu().numpy().reshape(hit_o.shape), hit_o) and \
            np.array_equal(a[2].cpu().numpy().reshape(tmin_o.shape).view(np.int32), tmin_o.view(np.int32))
        checked_oracle = globals().get('checked_oracle', 0) + 1
    if not ok:
        bad += 1
        print('MISMATCH seed', seed, 'N', N, 'n', n, int((a[1] != b[1]).sum()), 'rays')
print('pre-filter soak: %d scenes (seeds %d..%d), mismatching scenes: %d%s' % (
    count, first, first + count - 1, bad,
    ' (%d of them also compared with the C oracle: hit_index and tmin bit for bit)' % globals().get('checked_oracle', 0)
    if os.environ.get('ORACLE') else ''))
