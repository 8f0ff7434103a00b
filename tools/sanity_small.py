"""Smallest run that touches every kernel and code path (for compute-sanitizer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import oracle_c as oc, scenes
from reversible_raytracer_b200 import render as R
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from helpers import to_device

dev = torch.device('cuda:0')
cases = [scenes.match_mirror(n=20), scenes.orbit((3.8, -8.1, 32), 0, n=16), scenes.test_balls(n=12),
         scenes.stress(n=12, num_objects=1100), scenes.stress(n=10, num_objects=9, samples=3),
         scenes.stress(n=9, num_objects=7, samples=1, general=True)]
for spec in cases:
    ps = oc.PackedScene.from_spec(spec, camera_grad=1)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, dev)
    img, hit, tmin = R.render_forward(cfg, ot, w2o, mat, light, cam, jit, want_hit=True, want_tmin=True)
    g = R.render_backward(cfg, ot, w2o, mat, light, cam, torch.ones_like(img), hit, jit)
    g2 = R.render_backward(cfg, ot, w2o, mat, light, cam, torch.ones_like(img), None, jit)
    if cfg.samples in (1, 2, 4, 8):
        R.render_fused_mse(cfg, ot, w2o, mat, light, cam, torch.zeros_like(img), None, jit, want_image=True, want_hit=True)
    torch.cuda.synchronize()
from reversible_raytracer_b200.scene import *
from reversible_raytracer_b200.shader import *
c = torch.tensor([0.1, 0.2, 4.0], device=dev, requires_grad=True)
sc = Scene([Sphere(translate(c) * rotate(30, (0, 0, 1)) * scale((1, 2, 1.5)), Material((0.2, 0.9, 0.4), 0.3, 0.7, 0.5, 50.))],
           [Light((-1., -1., 2.), (0.961, 1., 0.87))], Camera(24, 24), PhongShader())
sc.build().sum().backward()
torch.cuda.synchronize()
print('sanity ok', c.grad.tolist())
