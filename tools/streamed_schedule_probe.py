"""StreamedFusedMSE step time for explicit slab-height schedules (graded ends, different middle heights),
C5 full image and the 512-row slab of one of 8 GPUs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reversible_raytracer_b200 import render as R, workloads as W, _native as nat
from tools.latency import timeit
dev = torch.device('cuda')
tb = W.stress_tables(1024); tt = W.stress_tables(1024, centre_noise=0.05)
t = lambda a: torch.from_numpy(a).to(dev)
args = (t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']), t(tb['camera']))


def graded(rows, mid, ramp):
    left = rows - 2 * sum(ramp)
    hs = list(ramp) + [mid] * (left // mid) + ([left % mid] if left % mid else []) + list(ramp[::-1])
    return hs


for rows, cases in ((4096, [('auto', None), ('mid128 r16', (128, [16, 32, 64])), ('mid192 r16', (192, [16, 32, 64, 128])),
                           ('mid256 r16', (256, [16, 32, 64, 128])), ('mid384 r16', (384, [16, 32, 64, 128, 256])),
                           ('mid256 r8', (256, [8, 16, 32, 64, 128])), ('mid512 r16', (512, [16, 32, 64, 128, 256]))]),
                    (512, [('auto', None), ('mid64 r4', (64, [4, 8, 16, 32])), ('mid96 r8', (96, [8, 16, 32])), ('mid128 r8', (128, [8, 16, 32, 64]))])):
    cfg = R.RenderConfig(n=4096, samples=4, shader=nat.SHADER_PHONG, transpose=1, seed=4321,
                         row_begin=0 if rows == 4096 else 1024, row_count=0 if rows == 4096 else rows)
    target = R.render_forward(cfg, args[0], t(tt['w2o']), *args[2:], None, want_hit=False)[0]
    pin_t = target.cpu().pin_memory(); pin_i = torch.empty_like(pin_t).pin_memory()
    ms = timeit(lambda: R.render_fused_mse(cfg, *args, target, want_image=True), warm=2, iters=8) / 1e3
    print('rows=%d resident single launch: %.3f ms' % (rows, ms))
    for name, spec in cases:
        st = R.StreamedFusedMSE(cfg, 1024, dev) if spec is None else R.StreamedFusedMSE(cfg, 1024, dev, heights=graded(rows, *spec))
        def step():
            l, g = st(*args, pin_t, pin_i)
            torch.cuda.current_stream().synchronize()
        ms = timeit(step, warm=2, iters=10) / 1e3
        print('   %-12s %2d slabs: %.3f ms' % (name, len(st.bounds), ms))
