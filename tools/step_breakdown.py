"""Where an optimise step of the small reference configs (C1, C3) spends its time:
host launch cost vs GPU time of the captured graph vs the loss read-back."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.latency import c1, c3, timeit


def breakdown(name, train):
    for _ in range(6):
        train()
    st = train.state
    g = st['graph']
    if g is None:
        print(name, 'no graph captured'); return
    torch.cuda.synchronize()
    n = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # GPU time of back-to-back replays (launch-overlapped)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print('%s: back-to-back replay %.1f us/step (GPU-side throughput)' % (name, e0.elapsed_time(e1) * 1e3 / n))
    # isolated replay GPU duration
    tot = 0.0
    for _ in range(n):
        torch.cuda.synchronize()
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    print('%s: isolated replay GPU duration %.1f us' % (name, tot * 1e3 / n))
    t0 = time.perf_counter()
    for _ in range(n):
        g.replay()
    t1 = time.perf_counter(); torch.cuda.synchronize()
    print('%s: host cost of replay() %.1f us' % (name, (t1 - t0) * 1e6 / n))
    t0 = time.perf_counter()
    for _ in range(n):
        g.replay(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    print('%s: replay + sync %.1f us' % (name, (t1 - t0) * 1e6 / n))
    print('%s: full train() %.1f us' % (name, timeit(train, warm=5, iters=n)))


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if which in ('all', 'c1'):
        breakdown('C1', c1()[0])
    if which in ('all', 'c3'):
        breakdown('C3 fused', c3('graph')[0])
