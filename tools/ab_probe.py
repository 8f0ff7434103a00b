"""A/B probes on the B200 box (exploration helper, not part of the product)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dataclasses import replace
from reversible_raytracer_b200 import render as R, workloads as W, _native as nat
from tools.latency import timeit, c4

dev = torch.device('cuda')
which = sys.argv[1] if len(sys.argv) > 1 else 'rays'
if which == 'rays':
    tb = W.stress_tables(1024)
    t = lambda a: torch.from_numpy(a).to(dev)
    args = (t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']), t(tb['camera']))
    cfg = R.RenderConfig(n=4096, samples=4, shader=nat.SHADER_PHONG, transpose=1, seed=4321)
    target = R.render_forward(cfg, *args, None, want_hit=False)[0]
    for lim in (1024, 4096):
        R.BASE_RAYS_MAX_N = lim
        us = timeit(lambda: R.render_fused_mse(cfg, *args, target, want_image=True), warm=2, iters=6)
        print('BASE_RAYS_MAX_N', lim, 'C5 fused ms', us / 1e3)
    img2 = R.render_forward(cfg, *args, None, want_hit=False)[0]
    print('table == in-kernel grid bitwise:', torch.equal(img2, target))
elif which == 'c4':
    fn, rays = c4(256)
    us = timeit(fn, warm=5, iters=200)
    print('C4 %.1f us  %.0f Mrays/s  RRT_SMALL_MAX_RAYS=%s' % (us, rays / us, os.environ.get('RRT_SMALL_MAX_RAYS')))
