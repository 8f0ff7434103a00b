"""Where does the per-step overhead after the render kernel go at N>1? (diagnostic)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
local = int(os.environ.get('LOCAL_RANK', '0')); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local); dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
G = 19477
grad = torch.randn(G, device=dev); loss = torch.zeros((), dtype=torch.float64, device=dev)
red64 = torch.zeros(G + 2, dtype=torch.float64, device=dev); red32 = torch.zeros(G + 2, dtype=torch.float32, device=dev)
busy = torch.empty(64 << 20, dtype=torch.float32, device=dev)
def ev(): return torch.cuda.Event(enable_timing=True)
for name, fn in [('copy+allreduce f64', lambda: (red64[:G].copy_(grad), red64[G].copy_(loss), dist.all_reduce(red64))),
                 ('allreduce f64 only', lambda: dist.all_reduce(red64)),
                 ('allreduce f32 only', lambda: dist.all_reduce(red32)),
                 ('allreduce f32 after 3ms busy', lambda: (busy.mul_(1.0001), busy.mul_(1.0001), busy.mul_(1.0001), busy.mul_(1.0001), busy.mul_(1.0001), busy.mul_(1.0001), busy.mul_(1.0001), busy.mul_(1.0001), dist.all_reduce(red32)))]:
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    tot = 0.0
    for _ in range(20):
        a, b = ev(), ev(); a.record(); fn(); b.record(); torch.cuda.synchronize(); tot += a.elapsed_time(b)
    if local == 0: print('%-32s %.1f us' % (name, tot / 20 * 1e3))
a, b = ev(), ev(); a.record()
for _ in range(8): busy.mul_(1.0001)
b.record(); torch.cuda.synchronize()
if local == 0: print('busy alone %.1f us' % (a.elapsed_time(b) * 1e3))
dist.destroy_process_group()
