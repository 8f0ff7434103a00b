"""single-workload driver for ncu: the C4 decoder batch (fused, d/d w2o only) with one thread mapping
(argv[1]: 1 = pixel per thread, 2 = ray per thread)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dataclasses import replace
from reversible_raytracer_b200 import render as R, workloads as W
mapping = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device('cuda')
tb, tt = W.orbit_tables(256), W.orbit_tables(256, centre_noise=0.5)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cfg = R.RenderConfig(n=64, samples=4, shader=tb['shader'], transpose=0, seed=7, camera_grad=0, geom_grad_only=1, pixel_threads=mapping)
args = (t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']), t(tb['camera']))
target, _, _ = R.render_forward(replace(cfg, pixel_threads=2), args[0], t(tt['w2o']), *args[2:], None, want_hit=False)
for _ in range(4):
    R.render_fused_mse(cfg, *args, target)
torch.cuda.synchronize()
