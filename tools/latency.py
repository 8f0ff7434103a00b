"""Optimise-step latency of the small reference configs (C1, C3) and throughput of the
batched autoencoder decoder workload (C4) -- exploration / reporting helper."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reversible_raytracer_b200 import render as R, workloads as W  # noqa: E402
from reversible_raytracer_b200.optimize import GDOptimizer  # noqa: E402
from reversible_raytracer_b200.scene import *  # noqa: E402,F401,F403
from reversible_raytracer_b200.shader import *  # noqa: E402,F401,F403


def timeit(fn, warm=5, iters=50):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e6


def c1(fused=False):
    """fused: 'graph' = the loss as a weight image through Scene.build_linear (fused kernel + CUDA graph),
    True = through Scene.linear_cost (whole step = one kernel launch)"""
    c1 = torch.tensor([-.5, -.5, 4.], device='cuda')
    c2 = torch.tensor([.5, .5, 4.], device='cuda')
    m1 = Material((0.2, 0.9, 0.4), 0.3, 0.7, 0.5, 50.)
    m2 = Material((0.87, 0.1, 0.507), 0.3, 0.9, 0.4, 50.)
    shapes = [Sphere(translate(c1), m1), Sphere(translate(c2) * rotate(90, (0, 0, 1)) * scale((1, 2, 1.5)), m2)]
    sc = Scene(shapes, [Light((-1., -1., 2.), (0.961, 1., 0.87))], Camera(128, 128), PhongShader())

    def loss():
        im = sc.build()
        return -im[90, 85].sum() - im[50, 90].sum()
    if fused:
        Wt = torch.zeros((128, 128, 3), device='cuda')
        Wt[90, 85] = -1.0
        Wt[50, 90] = -1.0
        loss = (lambda: sc.build_linear(Wt)) if fused == 'graph' else sc.linear_cost(Wt)
    train = GDOptimizer().optimize([c1, c2], loss, 0.0008, 0.1)
    return train, sc


def c2(fused=False):
    """BASELINE config C2 (test_balls.py:22-44): two spheres translate(p[:3]) * scale(p[3:]), DepthMapShader(6.1),
    32 x 32, loss = sum((X - image[:, :, 0]) ** 2) against the reference's 15.jpg, gradient descent on both
    6-vectors.  fused: the same cost through Scene.build_mse with channel weights (1, 0, 0); 'whole': through
    Scene.mse_cost (rrt_small_step_mse)."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    img = np.load(os.path.join(root, 'tests', 'golden', 'balls_15.npy')).astype(np.float32)
    if img.ndim == 3:
        img = img[:, :, 0]
    X = torch.tensor(img / 255.0, device='cuda')
    p1 = torch.tensor([-.4, -.3, 3., .5, .5, .5], device='cuda', requires_grad=True)   # (before slicing: views made earlier would not track it)
    p2 = torch.tensor([.4, .3, 3., .5, .5, .5], device='cuda', requires_grad=True)
    m = Material((0.2, 0.9, 0.4), 0.3, 0.7, 0.5, 50.)
    shapes = [Sphere(translate(p[:3]) * scale(p[3:]), m) for p in (p1, p2)]
    sc = Scene(shapes, [Light((-1., -1., 2.), (0.961, 1., 0.87))], Camera(32, 32), DepthMapShader(6.1))
    if fused:
        X3 = X[:, :, None].expand(32, 32, 3).contiguous()
        cost = lambda: sc.build_mse(X3, channel_weight=(1., 0., 0.), seed=15)
        if fused == 'whole':            # Scene.mse_cost: the whole step as one kernel launch
            cost = sc.mse_cost(X3, channel_weight=(1., 0., 0.), seed=15)
    else:
        cost = lambda: ((X - sc.build(seed=15)[:, :, 0]) ** 2).sum()
    return GDOptimizer().optimize([p1, p2], cost, 0.0001, 0.0), sc


def c3(fused):
    c1 = torch.tensor([-.5, -.5, 4.], device='cuda')
    c2 = torch.tensor([.5, .5, 4.], device='cuda')
    m1 = Material((0.2, 0.9, 0.4), 0.3, 0.7, 0.5, 50.)
    m2 = Material((0.87, 0.1, 0.507), 0.3, 0.9, 0.4, 50.)
    objs = [Sphere(translate(c1), m1), Sphere(translate(c2), m2), Square(translate((0, 0, 3)) * rotate(50, [0., 1., 0.]), m2)]
    sc = Scene(objs, [Light((-1., -1., 2.), (1., 0.87, 0.961))], Camera(128, 128), PhongShader())
    flipped = torch.flip(sc.build().detach(), dims=[1])
    cost = ((lambda: sc.build_mse(flipped)) if fused == 'graph' else sc.mse_cost(flipped)) if fused else \
        (lambda: ((sc.build() - flipped) ** 2).sum())
    return GDOptimizer().optimize([c1, c2], cost, 0.000008, 0.1), sc


def c4(num_scenes=256, geom_grad_only=1):
    """Decoder batch of the orbit autoencoders.  geom_grad_only=1: like every decoder of the reference
    (materials, light and camera direction are constants there) only d/d w2o is requested."""
    tb = W.orbit_tables(num_scenes)
    tt = W.orbit_tables(num_scenes, centre_noise=0.5)
    dev = torch.device('cuda')
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cfg = R.RenderConfig(n=64, samples=4, shader=tb['shader'], transpose=0, seed=7, camera_grad=0,
                         geom_grad_only=geom_grad_only)
    args = (t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']), t(tb['camera']))
    target, _, _ = R.render_forward(cfg, args[0], t(tt['w2o']), *args[2:], None, want_hit=False)
    return lambda: R.render_fused_mse(cfg, *args, target), 2 * num_scenes * 64 * 64 * 4


def c4_hits(num_scenes=256):
    """winning rays of the C4 decoder batch (for the credited flops)"""
    tb = W.orbit_tables(num_scenes)
    dev = torch.device('cuda')
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cfg = R.RenderConfig(n=64, samples=4, shader=tb['shader'], transpose=0, seed=7)
    _, hit, _ = R.render_forward(cfg, t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']), t(tb['camera']), None)
    return float((hit >= 0).sum())


def c4_closure():
    """One orbit scene pair through the drop-in API in the REFERENCE'S closure style
    (orbit_experiments/test_optimization.py:17-44: everything rebuilt inside the cost function on
    every call) + GDOptimizer.  -> (train, captured) ; captured() tells whether the step became a
    CUDA graph."""
    dev = torch.device('cuda')
    target = torch.rand((2, 64, 64, 3), device=dev)
    c = torch.tensor([3.0, -8.0, 32.0], device=dev)

    def scene(obj_param, cam_y, seed):
        material1 = Material((0.0, 0.9, 0.0), 0.3, 0.7, 0.5, 50.)
        material2 = Material((0.9, 0.0, 0.0), 0.3, 0.9, 0.4, 50.)
        center2 = np.asarray([0, 0, 48], dtype='float32')
        shapes = [Sphere(translate(obj_param) * scale((4, 4, 4)), material1),
                  Sphere(translate(center2) * scale((6, 6, 6)), material2)]
        light = Light((-0., -0., 1), (1., 1., 1.))
        camera = Camera(64, 64, translate((0, cam_y, 0)), np.asarray([0, 0, 1], dtype='float32'))
        return Scene(shapes, [light], camera, PhongShader(specular=False))

    def cost():
        return scene(c, 2.5, 5).build_mse(target[0], seed=5) + scene(c, -2.5, 6).build_mse(target[1], seed=6)
    train = GDOptimizer().optimize([c], cost, lr=1e-5)
    return train, (lambda: train.state['graph'] is not None)


if __name__ == '__main__':
    train, sc = c1()
    print('C1 optimize_brightness step: %.1f us' % timeit(train))
    print('C1 render only (scene.build): %.1f us' % timeit(lambda: sc.build()))
    for fused in (False, 'graph', True):
        train, sc = c3(fused)
        print('C3 match_mirror step (fused=%s%s): %.1f us' % (fused, ', whole-step kernel' if train.state.get('whole_step') else '', timeit(train)))
    fn, rays = c4()
    us = timeit(fn, iters=100)
    print('C4 orbit batch 256x2 fused fwd+mse+bwd: %.1f us  -> %.0f Mrays/s' % (us, rays / us))
