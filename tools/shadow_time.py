"""Cost of the opt-in hard-shadow pass on the stress scene (not a BASELINE config)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dataclasses import replace
from reversible_raytracer_b200 import render as R, workloads as W, _native as nat
from tools.latency import timeit
dev = torch.device('cuda')
for n, N in ((1024, 256), (4096, 1024)):
    tb = W.stress_tables(N); tt = W.stress_tables(N, centre_noise=0.05)
    t = lambda a: torch.from_numpy(a).to(dev)
    args = (t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']), t(tb['camera']))
    cfg = R.RenderConfig(n=n, samples=4, shader=nat.SHADER_PHONG, transpose=1, seed=4321)
    target = R.render_forward(cfg, args[0], t(tt['w2o']), *args[2:], None, want_hit=False)[0]
    for sh in (0, 2, 1):
        c = replace(cfg, shadows=sh)
        ms = timeit(lambda: R.render_fused_mse(c, *args, target, want_image=True), warm=1, iters=3) / 1e3
        _, hit, _ = R.render_forward(c, *args, None, want_hit=True)
        frac = float(((hit >= 0) & ((hit & nat.HIT_SHADOWED) != 0)).sum()) / max(1.0, float((hit >= 0).sum()))
        print('n=%d N=%d shadows=%d (2 = scalar pass only): fused %.2f ms, shadowed winners %.1f%%' % (n, N, sh, ms, 100 * frac))
