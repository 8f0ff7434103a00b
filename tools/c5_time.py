"""C5 / C5g fused step time with a realistic (perturbed) target, like bench.py: pre-filter sweep (default)
against the canonical sweep (RRT_FLAG_CANONICAL_SWEEP), and a bitwise comparison of their outputs.
RRT_B200_LIB=<other .so> A/Bs kernel builds.  usage: c5_time.py [n] [general]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dataclasses import replace
from reversible_raytracer_b200 import render as R, workloads as W, _native as nat
from tools.latency import timeit
dev = torch.device('cuda')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
general = len(sys.argv) > 2 and sys.argv[2] == 'general'
miss = len(sys.argv) > 2 and sys.argv[2] == 'miss'          # every sphere moved out of view: the pure sweep
tb = W.stress_tables(1024, general=general); tt = W.stress_tables(1024, general=general, centre_noise=0.05)
if miss:
    for t_ in (tb, tt):
        t_['w2o'].reshape(-1, 3, 4)[:, 0, 3] -= 1.0e4
t = lambda a: torch.from_numpy(a).to(dev)
args = (t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']), t(tb['camera']))
cfg = R.RenderConfig(n=n, samples=4, shader=nat.SHADER_PHONG, transpose=1, seed=4321)
target = R.render_forward(cfg, args[0], t(tt['w2o']), *args[2:], None, want_hit=False)[0]
lib = os.path.basename(os.environ.get('RRT_B200_LIB', 'default'))
outs = {}
for canon in (0, 1):
    c = replace(cfg, canonical_sweep=canon)
    outs[canon] = R.render_fused_mse(c, *args, target, want_image=True, want_hit=True)
    ms = timeit(lambda: R.render_fused_mse(c, *args, target, want_image=True), warm=2, iters=8) / 1e3
    mf = timeit(lambda: R.render_forward(c, *args, None, want_hit=False), warm=2, iters=8) / 1e3
    print('lib=%s %s n=%d canonical_sweep=%d: fused %.3f ms (%.0f Mrays/s), forward %.3f ms' %
          (lib, 'C5g' if general else ('C5-miss' if miss else 'C5'), n, canon, ms, n * n * 4 / ms / 1e3, mf))
a, b = outs[0], outs[1]
if not miss:
  print('pre-filter == canonical: hit %s, image %s, loss rel diff %.2e, grad rel diff %.2e' %
      (torch.equal(a[3], b[3]), torch.equal(a[2], b[2]), abs(float(a[0]) - float(b[0])) / abs(float(b[0])),
       float((a[1] - b[1]).abs().max() / b[1].abs().max())))
