"""Which piece of a reference-style loss closure breaks CUDA-graph capture? (exploration helper)"""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from reversible_raytracer_b200.scene import *
from reversible_raytracer_b200.shader import *
from reversible_raytracer_b200.optimize import GDOptimizer
cuda = torch.device('cuda')
theta = torch.tensor(0.7, device=cuda, requires_grad=True)
z32 = torch.tensor(32.0, device=cuda)
m1 = Material((0.0, 0.9, 0.0), 0.3, 0.7, 0.5, 50.)
m2 = Material((0.9, 0.0, 0.0), 0.3, 0.9, 0.4, 50.)
cams = [Camera(64, 64, translate((0, y, 0)), np.asarray([0, 0, 1], dtype='float32')) for y in (2.5, -2.5)]
scs = [Scene([], [Light((0., 0., 1.), (1., 1., 1.))], cam, PhongShader(specular=False)) for cam in cams]
target = torch.rand((2, 64, 64, 3), device=cuda)

def centre():
    return torch.stack([9 * torch.cos(theta), 9 * torch.sin(theta), z32])

def build(v):
    c = centre()
    sc = scs[v]
    sc.shapes = [Sphere(translate(c) * scale((4., 4., 4.)), m1), Sphere(translate((0, 0, 48)) * scale((6, 6, 6)), m2)]
    return sc.build(seed=7 + v)

def cost():
    return sum(((build(v) - target[v]) ** 2).sum() for v in range(2))

def grad_step():
    v = cost()
    return torch.autograd.grad(v, [theta])[0]

pieces = {'centre': centre, 'build0': lambda: build(0), 'cost': cost, 'grad': grad_step}
for name, fn in pieces.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, stream=side):
            try:
                fn()
            except Exception:
                print('--- inside capture of', name)
                traceback.print_exc()
                raise
        g.replay()
        torch.cuda.synchronize()
        print('capture ok:', name)
    except Exception as e:
        print('capture FAILED:', name, repr(e)[:300])
        torch.cuda.synchronize()
