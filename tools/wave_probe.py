"""C5 fused step on row slabs whose CTA count is an exact multiple of the resident CTA slots (148 SMs x 4 = 592;
a CTA is 64 x 4 pixels => 148 rows = 4 full waves): time vs waves => the per-launch constant (intercept) and the
time per wave (slope).  Graph-replayed (no host overhead)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from reversible_raytracer_b200 import render as R, workloads as W, _native as nat
from tools.latency import timeit
dev = torch.device('cuda')
tb = W.stress_tables(1024); tt = W.stress_tables(1024, centre_noise=0.05)
MISS, TICKET, FWD = int(os.environ.get('MISS', '0')), int(os.environ.get('TICKET', '1')), int(os.environ.get('FWD', '0'))
if MISS:                                   # every sphere out of view: pure sweep, no shading / reverse pass
    for t_ in (tb, tt):
        t_['w2o'].reshape(-1, 3, 4)[:, 0, 3] -= 1.0e4
print('MISS=%d TICKET=%d FWD=%d' % (MISS, TICKET, FWD))
t = lambda a: torch.from_numpy(a).to(dev)
args = (t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']), t(tb['camera']))
for rows in (4, 148, 296, 592, 1184):
    cfg = R.RenderConfig(n=4096, samples=4, shader=nat.SHADER_PHONG, transpose=1, seed=4321, row_begin=1024, row_count=rows,
                         use_ticket=TICKET)
    target = R.render_forward(cfg, args[0], t(tt['w2o']), *args[2:], None, want_hit=False)[0]
    fn = (lambda: R.render_forward(cfg, *args, None, want_hit=False)) if FWD else \
        (lambda: R.render_fused_mse(cfg, *args, target, want_image=True))
    g, s = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g, stream=s):
        fn()
    us = timeit(g.replay, warm=5, iters=40)
    ctas = 64 * ((rows + 3) // 4)
    print('rows=%4d  CTAs=%6d = %6.2f waves: %8.1f us  (%.1f us per wave)' % (rows, ctas, ctas / 592.0, us, us / (ctas / 592.0)))
