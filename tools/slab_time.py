"""Time the C5 fused step on row slabs of 4096/G rows (what each of G GPUs renders) for the
resident-CTA cap given by RRT_MAX_RESIDENT (experiment: last-wave quantisation)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reversible_raytracer_b200 import render as R, workloads as W, _native as nat
from tools.latency import timeit
dev = torch.device('cuda')
tb = W.stress_tables(1024); tt = W.stress_tables(1024, centre_noise=0.05)
t = lambda a: torch.from_numpy(a).to(dev)
args = (t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']), t(tb['camera']))
out = []
for G in (1, 2, 4, 8):
    rows = 4096 // G
    cfg = R.RenderConfig(n=4096, samples=4, shader=nat.SHADER_PHONG, transpose=1, seed=4321, row_begin=1024, row_count=rows)
    if G == 1:
        cfg = R.RenderConfig(n=4096, samples=4, shader=nat.SHADER_PHONG, transpose=1, seed=4321)
    target = R.render_forward(cfg, args[0], t(tt['w2o']), *args[2:], None, want_hit=False)[0]
    ms = timeit(lambda: R.render_fused_mse(cfg, *args, target, want_image=True), warm=2, iters=8) / 1e3
    out.append('G=%d %.3f ms (x%d = %.2f)' % (G, ms, G, ms * G))
print('RRT_MAX_RESIDENT=%s  ' % os.environ.get('RRT_MAX_RESIDENT', '5') + '  '.join(out))
