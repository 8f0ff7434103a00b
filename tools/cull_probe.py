import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dataclasses import replace
from reversible_raytracer_b200 import render as R, workloads as W, _native as nat
from tools.latency import timeit
dev = torch.device('cuda')
for general in (False, True):
    tb = W.stress_tables(1024, general=general)
    t = lambda a: torch.from_numpy(a).to(dev)
    args = (t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']), t(tb['camera']))
    cfg = R.RenderConfig(n=4096, samples=4, shader=nat.SHADER_PHONG, transpose=1, seed=4321)
    a = R.render_forward(cfg, *args, None, want_hit=True, want_tmin=True)
    b = R.render_forward(replace(cfg, cull=1), *args, None, want_hit=True, want_tmin=True)
    print('general', general, 'full image bitwise equal:', all(torch.equal(x, y) for x, y in zip(a, b)))
    target = a[0].clone(); del a, b
    for cull in (0, 1):
        c = replace(cfg, cull=cull)
        us = timeit(lambda: R.render_fused_mse(c, *args, target, want_image=True), warm=2, iters=5)
        print('  cull', cull, 'fused ms', us / 1e3, 'Mrays/s', 4096 * 4096 * 4 / us)
