"""Host-side cost of one eager Scene.build() (no CUDA graph): cProfile of 2000 calls."""
import os, sys, cProfile, pstats, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.latency import c1
train, sc = c1()
for _ in range(20):
    sc.build()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000):
    sc.build()
torch.cuda.synchronize()
print('Scene.build eager: %.1f us/call' % ((time.perf_counter() - t0) / 2000 * 1e6))
pr = cProfile.Profile()
pr.enable()
for _ in range(2000):
    sc.build()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(18)
print('\n'.join(s.getvalue().splitlines()[:40]))
