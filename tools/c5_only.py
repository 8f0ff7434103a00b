"""One C5 fused step (default sweep) for ncu captures.  usage: c5_only.py [canonical]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reversible_raytracer_b200 import render as R, workloads as W, _native as nat
dev = torch.device('cuda')
canon = int(len(sys.argv) > 1 and sys.argv[1] == 'canonical')
tb = W.stress_tables(1024); tt = W.stress_tables(1024, centre_noise=0.05)
t = lambda a: torch.from_numpy(a).to(dev)
args = (t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']), t(tb['camera']))
cfg = R.RenderConfig(n=4096, samples=4, shader=nat.SHADER_PHONG, transpose=1, seed=4321, canonical_sweep=canon)
target = R.render_forward(cfg, args[0], t(tt['w2o']), *args[2:], None, want_hit=False)[0]
for _ in range(3):
    loss, grad, img, _ = R.render_fused_mse(cfg, *args, target, want_image=True)
torch.cuda.synchronize()
print('loss', float(loss))
