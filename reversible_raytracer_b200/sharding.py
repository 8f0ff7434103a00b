"""Multi-GPU sharding of the render path (one process per GPU, torch.distributed).

Rays are independent in the forward (scene.py:27-50 has no cross-pixel term) and
the reverse pass is a SUM over rays of per-ray parameter gradients, so:
  * one big image  -> contiguous ROW SLABS, one per rank; every rank keeps its own
    image / target / hit-index slab, nothing is gathered;
  * a batch of scenes (autoencoder workloads) -> contiguous SCENE RANGES per rank;
  * the only exchange is ONE sum of the flat vector [gradient, loss] over ranks
    (N*19+21+1 numbers): on GPUs `PeerSum` -- our own kernel over NVLink peer memory
    (rrt_peer_allreduce: push to every peer, flags, sum in rank order) -- with an NCCL
    allreduce as the plain alternative (`allreduce_loss_grad`); gloo in the CPU tests.
"""
import ctypes as C

import torch
import torch.distributed as dist


def row_slab(n, world, rank):
    """(row_begin, row_count) of `rank` for an n-row image split into `world`
    contiguous slabs; the first n % world ranks get one extra row.  Every rank gets
    at least one row when world <= n."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError('bad world/rank')
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, base + (1 if rank < extra else 0)


def balanced_row_slabs(row_cost, world, align=4):
    """Contiguous row slabs of (nearly) EQUAL TOTAL COST instead of equal height: `row_cost[a]` is the estimated
    cost of image row a (e.g. rays + k * winning rays, from a previous render's hit mask -- the sweep costs the
    same everywhere, shading and the reverse pass follow where the objects are).  Slab boundaries are multiples
    of `align` rows (the render kernel's CTAs are 4 rows high), every rank gets at least `align` rows (or one
    row when there are fewer rows than that), the slabs cover [0, n) in rank order.  Deterministic: every rank
    computes the same partition from the same costs.  -> list of (row_begin, row_count) per rank."""
    import numpy as np
    cost = np.asarray(row_cost, dtype=np.float64).reshape(-1)
    n = int(cost.size)
    if world < 1 or n < world:
        raise ValueError('need at least one row per rank')
    if not np.all(np.isfinite(cost)) or np.any(cost < 0) or cost.sum() <= 0:
        cost = np.ones(n)
    align = max(1, int(align))
    if n < world * align:
        align = 1
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    bounds = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        b = int(np.searchsorted(cum, target))                 # first boundary whose prefix cost reaches the target
        if b > 0 and abs(cum[b - 1] - target) <= abs(cum[b] - target):
            b -= 1
        b = int(round(b / align)) * align
        lo, hi = bounds[-1] + align, n - (world - r) * align  # leave room for the ranks before and after
        bounds.append(min(max(b, lo), hi))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1] - bounds[r]) for r in range(world)]


def scene_range(num_scenes, world, rank):
    """(first, count) of the scenes of `rank` (same split rule as row_slab)."""
    return row_slab(num_scenes, world, rank)


def allreduce_loss_grad(loss, grad, group=None):
    """Sum (loss, grad) over ranks with ONE collective.  loss: scalar or [B] tensor,
    grad: [..., G] tensor.  Returns (loss, grad) as float64 / grad.dtype.  No-op when
    torch.distributed is not initialised (single GPU)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return loss, grad
    flat = torch.cat([grad.reshape(-1).to(torch.float64), loss.reshape(-1).to(torch.float64)])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    g = flat[:grad.numel()].reshape(grad.shape).to(grad.dtype)
    l = flat[grad.numel():].reshape(loss.shape)
    return l, g


class PeerSum(object):
    """Sum of [grad (n float32) | loss (nloss float64)] over the ranks of one box by
    rrt_peer_allreduce (include/rrt_b200.h): one kernel per rank over NVLink peer memory,
    no NCCL call and no packing copies; every rank gets bit-identical float64 sums (fixed
    rank order).  Buffers come from torch's symmetric-memory allocator (cuMem + handle
    exchange through the process group's store); construction is collective."""

    def __init__(self, n, nloss=1, device=None, group=None):
        import torch.distributed._symmetric_memory as symm
        from . import _native as nat
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError('PeerSum needs an initialised torch.distributed process group')
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.n, self.nloss = int(n), int(nloss)
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        L = nat.lib()
        sig = (int(L.rrt_peer_signal_bytes()) + 255) // 256 * 256
        nbytes = sig + int(L.rrt_peer_buffer_bytes(self.n, self.nloss, self.world))
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.buf.zero_()
        torch.cuda.synchronize(self.device)
        self.handle = symm.rendezvous(self.buf, self.group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.peer_sig = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        self.peer_buf = torch.tensor([p + sig for p in ptrs], dtype=torch.int64, device=self.device)
        self.out = torch.empty(self.n + self.nloss, dtype=torch.float64, device=self.device)
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)               # every rank's flag area is zeroed before anyone pushes

    def __call__(self, grad, loss):
        """grad: contiguous float32 [n], loss: float64 [nloss] -> (loss_sum [nloss], grad_sum [n]) float64
        views of an internal buffer (valid until the next call)."""
        from . import _native as nat
        if grad.dtype != torch.float32 or grad.numel() != self.n or not grad.is_contiguous():
            raise ValueError('grad must be a contiguous float32 tensor of %d elements' % self.n)
        loss = loss.reshape(-1)
        if loss.dtype != torch.float64 or loss.numel() != self.nloss:
            raise ValueError('loss must be float64 with %d elements' % self.nloss)
        with torch.cuda.device(self.device):
            rc = nat.lib().rrt_peer_allreduce(grad.data_ptr(), loss.data_ptr(), self.n, self.nloss,
                                              self.peer_buf.data_ptr(), self.peer_sig.data_ptr(), self.rank, self.world,
                                              self.out.data_ptr(),
                                              C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        nat.check(rc, 'rrt_peer_allreduce')
        return self.out[self.n:], self.out[:self.n]


def render_fused_mse_sharded(cfg, obj_type, w2o, material, light, camera, target_slab, channel_weight=None,
                             jitter_slab=None, group=None, peer_sum=None, slab=None):
    """Row-slab sharded fused forward + MSE + reverse pass: this rank renders rows
    row_slab(n, world, rank) -- or `slab` = (row_begin, row_count), e.g. this rank's entry of
    balanced_row_slabs(...) -- against its resident `target_slab`, then the gradient
    vector and loss are summed over ranks.  Returns (loss, grad) identical on every rank."""
    from . import render as R
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    rb, rc = slab if slab is not None else row_slab(cfg.n, world, rank)
    loss, grad, _, _ = R.render_fused_mse(cfg.slab(rb, rc), obj_type, w2o, material, light, camera, target_slab,
                                          channel_weight, jitter_slab)
    if peer_sum is not None:               # our kernel over NVLink peer memory instead of NCCL
        l, g = peer_sum(grad.reshape(-1), loss.reshape(-1))
        return l.reshape(loss.shape), g.reshape(grad.shape).to(grad.dtype)
    return allreduce_loss_grad(loss, grad, group)
