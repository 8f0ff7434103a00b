"""Multi-GPU sharding of the render path (one process per GPU, torch.distributed).

Rays are independent in the forward (scene.py:27-50 has no cross-pixel term) and
the reverse pass is a SUM over rays of per-ray parameter gradients, so:
  * one big image  -> contiguous ROW SLABS, one per rank; every rank keeps its own
    image / target / hit-index slab, nothing is gathered;
  * a batch of scenes (autoencoder workloads) -> contiguous SCENE RANGES per rank;
  * the only exchange is ONE allreduce(sum) of the flat vector [gradient, loss]
    (N*19+21+1 numbers; NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
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


def render_fused_mse_sharded(cfg, obj_type, w2o, material, light, camera, target_slab, channel_weight=None,
                             jitter_slab=None, group=None):
    """Row-slab sharded fused forward + MSE + reverse pass: this rank renders rows
    row_slab(n, world, rank) against its resident `target_slab`, then the gradient
    vector and loss are summed over ranks.  Returns (loss, grad) identical on every rank."""
    from . import render as R
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    rb, rc = row_slab(cfg.n, world, rank)
    loss, grad, _, _ = R.render_fused_mse(cfg.slab(rb, rc), obj_type, w2o, material, light, camera, target_slab,
                                          channel_weight, jitter_slab)
    return allreduce_loss_grad(loss, grad, group)
