"""StreamedFusedMSE (host target in, image out) step time vs the number of row slabs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reversible_raytracer_b200 import render as R, workloads as W, _native as nat
from tools.latency import timeit
dev = torch.device('cuda')
tb = W.stress_tables(1024); tt = W.stress_tables(1024, centre_noise=0.05)
t = lambda a: torch.from_numpy(a).to(dev)
args = (t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']), t(tb['camera']))
cfg = R.RenderConfig(n=4096, samples=4, shader=nat.SHADER_PHONG, transpose=1, seed=4321)
target = R.render_forward(cfg, args[0], t(tt['w2o']), *args[2:], None, want_hit=False)[0]
pin_t = target.cpu().pin_memory(); pin_i = torch.empty_like(pin_t).pin_memory()
ms = timeit(lambda: R.render_fused_mse(cfg, *args, target, want_image=True), warm=2, iters=6) / 1e3
print('resident single launch: %.3f ms' % ms)
for slabs in (4, 8, 12, 16, 24, 32):
    st = R.StreamedFusedMSE(cfg, 1024, dev, slabs=slabs)
    def step():
        l, g = st(*args, pin_t, pin_i)
        torch.cuda.current_stream().synchronize()
    ms = timeit(step, warm=2, iters=6) / 1e3
    print('slabs=%2d: %.3f ms' % (slabs, ms))
