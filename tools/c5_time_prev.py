"""Times the C5 fused step with whatever library RRT_B200_LIB points at, using only the C-ABI
fields both the previous and the current build know (A/B across builds in ONE gpurun call)."""
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
cfg = R.RenderConfig(n=4096, samples=4, shader=nat.SHADER_PHONG, transpose=1, seed=4321, use_records=int(os.environ.get('REC', '0')))
target = R.render_forward(cfg, args[0], t(tt['w2o']), *args[2:], None, want_hit=False)[0]
ms = [timeit(lambda: R.render_fused_mse(cfg, *args, target, want_image=True), warm=2, iters=8) / 1e3 for _ in range(2)]
print(os.path.basename(os.environ.get('RRT_B200_LIB', 'current')), 'REC=%s' % os.environ.get('REC', '0'), ' '.join('%.3f' % m for m in ms))
