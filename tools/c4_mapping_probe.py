"""C4 orbit batch (N scenes x 2 views, 64x64, S=4): thread mapping of the small-scene kernel
(1 = one pixel per thread, 2 = one ray per thread), fused and forward, graph-replayed GPU time.
RRT_B200_LIB=<other .so> A/Bs kernel builds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dataclasses import replace
from reversible_raytracer_b200 import render as R, workloads as W
from tools.latency import timeit

dev = torch.device('cuda')
for scenes_ in [int(a) for a in sys.argv[1:]] or [256, 32]:
    tb, tt = W.orbit_tables(scenes_), W.orbit_tables(scenes_, centre_noise=0.5)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cfg = R.RenderConfig(n=64, samples=4, shader=tb['shader'], transpose=0, seed=7, camera_grad=0)
    args = (t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']), t(tb['camera']))
    target, _, _ = R.render_forward(cfg, args[0], t(tt['w2o']), *args[2:], None, want_hit=False)
    rays = 2 * scenes_ * 64 * 64 * 4
    _, hit_all, _ = R.render_forward(cfg, *args, None, want_hit=True)
    dl = torch.randn_like(target)
    for what in ('fused geom', 'fused all', 'forward', 'backward stored', 'backward resweep'):
        for mapping in (1, 2):
            c = replace(cfg, geom_grad_only=int(what == 'fused geom'), pixel_threads=mapping)
            if what == 'forward':
                fn = lambda: R.render_forward(c, *args, None, want_hit=False)
            elif what.startswith('backward'):
                fn = lambda: R.render_backward(c, *args, dl, hit_all if what.endswith('stored') else None)
            else:
                fn = lambda: R.render_fused_mse(c, *args, target)
            g, s = torch.cuda.CUDAGraph(), torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                fn()
            torch.cuda.current_stream().wait_stream(s)
            with torch.cuda.graph(g, stream=s):
                fn()
            ug = timeit(g.replay, warm=10, iters=300)
            print('lib=%s scenes=%d %-10s mapping=%s: %.1f us graph replay -> %.0f Mrays/s' %
                  (os.path.basename(os.environ.get('RRT_B200_LIB', 'default')), scenes_, what,
                   {1: 'pixel', 2: 'ray'}[mapping], ug, rays / ug))
