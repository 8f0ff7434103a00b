"""PeerSum (rrt_peer_allreduce, our kernel over NVLink peer memory) vs NCCL allreduce:
correctness on random vectors over many epochs, then latency of both.  Run under torchrun."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from reversible_raytracer_b200 import sharding as Sh

local = int(os.environ.get('LOCAL_RANK', '0')); world = int(os.environ['WORLD_SIZE']); rank = int(os.environ['RANK'])
torch.cuda.set_device(local); dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
G = 1024 * 19 + 21
ps = Sh.PeerSum(G, 1, dev)
gen = torch.Generator(device=dev); gen.manual_seed(100 + rank)
bad = 0
for it in range(200):
    grad = torch.randn(G, device=dev, generator=gen)
    loss = torch.rand(1, device=dev, generator=gen, dtype=torch.float64)
    if it % 7 == rank % 7:                 # desynchronise the ranks: somebody is always late
        torch.cuda._sleep(2_000_000)
    l, g = ps(grad, loss)
    ref = torch.cat([grad.double(), loss]); dist.all_reduce(ref)
    if not (torch.allclose(g, ref[:G], rtol=1e-13, atol=1e-13) and torch.allclose(l, ref[G:], rtol=1e-13)):
        bad += 1
# identical bits on every rank
chk = torch.cat([g, l]).clone(); lst = [torch.empty_like(chk) for _ in range(world)]; dist.all_gather(lst, chk)
same = all(torch.equal(lst[0], t) for t in lst)
def ev(): return torch.cuda.Event(enable_timing=True)
red = torch.zeros(G + 2, dtype=torch.float64, device=dev)
def nccl_path():
    red[:G] = grad; red[G] = loss[0]; dist.all_reduce(red)
res = {}
for name, fn in (('peer kernel', lambda: ps(grad, loss)), ('NCCL (2 pack copies + allreduce f64)', nccl_path)):
    for _ in range(10): fn()
    torch.cuda.synchronize(); dist.barrier()
    tot = 0.0
    for _ in range(50):
        a, b = ev(), ev(); a.record(); fn(); b.record(); torch.cuda.synchronize(); tot += a.elapsed_time(b)
    t = torch.tensor([tot / 50 * 1e3], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); res[name] = float(t)
if rank == 0:
    print('PeerSum: world %d, %d epochs, mismatches vs NCCL: %d, identical bits on all ranks: %s' % (world, 200, bad, same))
    for k, v in res.items(): print('  %-40s %.1f us (max over ranks)' % (k, v))
dist.barrier(); dist.destroy_process_group()
