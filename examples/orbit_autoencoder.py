"""Batched B200 port of the orbit autoencoder experiment
(orbit_experiments/autoencoder_2ly.py + test_optimization.py): a small MLP encoder
maps each camera view to a sphere centre; the DECODER is the differentiable ray
tracer; cost = sum of squared errors of both views (autoencoder_2ly.py:87-91).

The reference trains one scene pair per compiled call ("TODO remake it for batch",
autoencoder_2ly.py:116); here a whole batch of scene pairs is one fused
forward + loss + reverse-pass kernel launch, and the encoder is stock PyTorch.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reversible_raytracer_b200 import render as R, workloads as W  # noqa: E402


class Encoder(torch.nn.Module):
    """autoencoder_2ly.py:62-80: tanh(X W0 + b) -> tanh(. W1 + b) -> centre = h2 Cw + cbias, relu on z."""

    def __init__(self, n_visible, h1=600, h2=30):
        super().__init__()
        self.l1, self.l2, self.cap = torch.nn.Linear(n_visible, h1), torch.nn.Linear(h1, h2), torch.nn.Linear(h2, 3)
        with torch.no_grad():
            self.cap.weight.mul_(0.05)
            self.cap.bias.copy_(torch.tensor([0., 0., 32.]))

    def forward(self, x):
        c = self.cap(torch.tanh(self.l2(torch.tanh(self.l1(x)))))
        return torch.cat([c[..., :2], torch.relu(c[..., 2:])], dim=-1)


def make_trainer(num_scenes=256, lr=2e-8, n=64, seed=1234, dev=None, world=1, rank=0, graph=False):
    """`graph=True`: after three eager steps the WHOLE training step (encoder forward, fused render
    launch, encoder backward, the NCCL allreduce, SGD update) is captured into ONE CUDA graph and
    `step()` replays it -- the counterpart of the reference's single compiled `train` function
    (orbit_experiments/optimize.py:75-93); falls back to eager stepping if the capture fails.
    -> (step, info): `step()` runs ONE training step of the batched orbit autoencoder on this
    rank's scene range -- encoder forward (stock PyTorch), decoder = ONE fused render + squared
    error + reverse-pass launch (only d/d w2o is requested: materials, light and camera are
    constants, autoencoder_2ly.py:82-91), encoder backward, ONE flat NCCL allreduce of the
    encoder-weight gradients + loss when world > 1, SGD update -- and returns
    the loss tensor (global sum).  `info` has the sizes."""
    import torch.distributed as dist
    from reversible_raytracer_b200 import sharding
    tb = W.orbit_tables(num_scenes, seed=seed)
    first, count = sharding.scene_range(num_scenes, world, rank)          # this rank's scenes (x2 views)
    for key in ('w2o', 'camera'):
        tb[key] = tb[key][2 * first:2 * (first + count)]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cfg = R.RenderConfig(n=n, samples=4, shader=tb['shader'], transpose=0, seed=11, scene_begin=2 * first, geom_grad_only=1)
    obj_type, material, light, camera = t(tb['obj_type']), t(tb['material']), t(tb['light']), t(tb['camera'])
    B = 2 * count
    # data: the scenes rendered at their true centres (planet_orbit.py), as uint8 like the dataset
    X, _, _ = R.render_forward(cfg, obj_type, t(tb['w2o']), material, light, camera, None, want_hit=False)
    X = (X * 255).to(torch.uint8).float() / 255.0                                        # [B,n,n,3]
    torch.manual_seed(0)                                     # identical replicas on every rank
    enc = Encoder(n * n * 3).to(dev)
    opt = torch.optim.SGD(enc.parameters(), lr=lr)
    fixed = torch.tensor([0., 0., 48.], device=dev).expand(B, 3)
    scales = torch.tensor([[4., 4., 4.], [6., 6., 6.]], device=dev).expand(B, 2, 3)
    params = list(enc.parameters())
    nparam = sum(p.numel() for p in params)
    flat = torch.empty(nparam + 1, dtype=torch.float32, device=dev)

    def step():
        centres = enc(X.reshape(B, -1))                                                  # [B,3] (one per view)
        w2o = R.w2o_translate_scale(torch.stack([centres, fixed], dim=1), scales)        # [B,2,12]
        loss = R.render_fused_mse_loss(cfg, obj_type, w2o, material, light, camera, X).sum()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if world > 1:                                        # one flat allreduce: [grads..., loss]
            torch.cat([p.grad.reshape(-1) for p in params] + [loss.detach().reshape(1)], out=flat)
            dist.all_reduce(flat)
            off = 0
            for p in params:
                p.grad.copy_(flat[off:off + p.numel()].reshape(p.shape))
                off += p.numel()
            loss = flat[-1]
        opt.step()
        return loss.detach()
    info = dict(scenes_per_rank=count, views_per_rank=B, encoder_parameters=nparam,
                allreduce_bytes=(nparam + 1) * 4 if world > 1 else 0, rays_per_rank=B * n * n * 4, cuda_graph=False)
    if not graph:
        return step, info
    # ---- whole step as one CUDA graph (warm-up on a side stream, like torch's whole-network capture recipe)
    eager = step
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            eager()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    ok = torch.ones(1, device=dev)
    g, out = torch.cuda.CUDAGraph(), None
    try:
        with torch.cuda.graph(g):
            out = eager()
    except Exception as e:          # noqa: BLE001
        sys.stderr.write('orbit_autoencoder: capture of the training step failed (%r); stepping eagerly\n' % (e,))
        ok.zero_()
        torch.cuda.synchronize(dev)
    if world > 1:                   # all ranks replay, or none does (the graph contains a collective)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if float(ok) < 1:
        return eager, info
    info['cuda_graph'] = True

    def replay():
        g.replay()
        return out
    return replay, info


def main(num_scenes=256, steps=30, lr=2e-8, n=64, seed=1234, verbose=True, graph=False):
    """Single GPU, or `torchrun --nproc-per-node G examples/orbit_autoencoder.py`: the scene
    batch is sharded across ranks (sharding.scene_range), every rank renders and
    back-propagates its own scenes, and ONE flat allreduce sums the encoder-weight
    gradients."""
    import torch.distributed as dist
    world, rank = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group('nccl', device_id=dev)
    step, _ = make_trainer(num_scenes, lr, n, seed, dev, world, rank, graph=graph)
    losses = []
    for k in range(steps):
        losses.append(float(step()))
        if verbose and rank == 0:
            print('step %d cost %.3f' % (k, losses[-1]))
    return losses


if __name__ == '__main__':
    main(graph='--graph' in sys.argv)
