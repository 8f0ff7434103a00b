"""StreamedFusedMSE (host target in, image out) step time vs the number of row slabs, for the
full image (1 GPU) and for the 512-row slab one of 8 GPUs renders."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reversible_raytracer_b200 import render as R, workloads as W, _native as nat
from tools.latency import timeit
dev = torch.device('cuda')
tb = W.stress_tables(1024); tt = W.stress_tables(1024, centre_noise=0.05)
t = lambda a: torch.from_numpy(a).to(dev)
args = (t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']), t(tb['camera']))
for rows, slab_list in ((512, (1, 2, 3, 4, 6, 8, 12, 16)), (1024, (4, 6, 8, 12, 16)), (4096, (16, 24, 32))):
    cfg = R.RenderConfig(n=4096, samples=4, shader=nat.SHADER_PHONG, transpose=1, seed=4321,
                         row_begin=0 if rows == 4096 else 1024, row_count=0 if rows == 4096 else rows)
    target = R.render_forward(cfg, args[0], t(tt['w2o']), *args[2:], None, want_hit=False)[0]
    pin_t = target.cpu().pin_memory(); pin_i = torch.empty_like(pin_t).pin_memory()
    ms = timeit(lambda: R.render_fused_mse(cfg, *args, target, want_image=True), warm=2, iters=8) / 1e3
    print('rows=%d resident single launch: %.3f ms' % (rows, ms))
    for slabs in slab_list:
        st = R.StreamedFusedMSE(cfg, 1024, dev, slabs=slabs)
        def step():
            l, g = st(*args, pin_t, pin_i)
            torch.cuda.current_stream().synchronize()
        ms = timeit(step, warm=2, iters=8) / 1e3
        print('   slabs=%2d (%d rows each): %.3f ms' % (slabs, st.bounds[0][1], ms))
