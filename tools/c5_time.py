"""C5 fused step time with a realistic (perturbed) target, like bench.py; A/B of the record table."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dataclasses import replace
from reversible_raytracer_b200 import render as R, workloads as W, _native as nat
from tools.latency import timeit
dev = torch.device('cuda')
tb = W.stress_tables(1024); tt = W.stress_tables(1024, centre_noise=0.05)
t = lambda a: torch.from_numpy(a).to(dev)
args = (t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']), t(tb['camera']))
cfg = R.RenderConfig(n=4096, samples=4, shader=nat.SHADER_PHONG, transpose=1, seed=4321)
target = R.render_forward(cfg, args[0], t(tt['w2o']), *args[2:], None, want_hit=False)[0]
for rep in range(2):
    for rec in (1, 0):
        c = replace(cfg, use_records=rec)
        ms = timeit(lambda: R.render_fused_mse(c, *args, target, want_image=True), warm=2, iters=8) / 1e3
        print('use_records=%d C5 fused ms: %.3f' % (rec, ms))
