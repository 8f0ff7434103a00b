"""A minimal orbit autoencoder in the reference's shape (orbit_experiments/autoencoder_2ly.py:62-91 +
test_optimization.py:17-44): a linear encoder maps the two camera views of a sample to a sphere centre, the
decoder rebuilds materials, shapes, light, cameras and Scene inside cost() on every call."""
import numpy as np
import torch

from reversible_raytracer_b200.scene import Camera, Light, Material, Scene, Sphere, scale, translate
from reversible_raytracer_b200.shader import PhongShader


class OrbitAE(object):
    def __init__(self, n, device, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.n = n
        self.W = (torch.randn(2 * n * n * 3, 3, generator=g) * 1e-4).to(device)
        self.b = torch.tensor([1.0, -2.0, 30.0], device=device)
        self.params = [self.W, self.b]

    def scene(self, centre, cam_y, seed):
        material1 = Material((0.0, 0.9, 0.0), 0.3, 0.7, 0.5, 50.)
        material2 = Material((0.9, 0.0, 0.0), 0.3, 0.9, 0.4, 50.)
        shapes = [Sphere(translate(centre) * scale((4, 4, 4)), material1),
                  Sphere(translate(np.asarray([0, 0, 48], dtype='float32')) * scale((6, 6, 6)), material2)]
        camera = Camera(self.n, self.n, translate((0, cam_y, 0)), np.asarray([0, 0, 1], dtype='float32'))
        return Scene(shapes, [Light((-0., -0., 1), (1., 1., 1.))], camera, PhongShader(specular=False))

    def cost(self, Xl, Xr):
        centre = torch.cat([Xl.reshape(-1), Xr.reshape(-1)]) @ self.W + self.b
        return self.scene(centre, 2.5, 5).build_mse(Xl, seed=5) + self.scene(centre, -2.5, 6).build_mse(Xr, seed=6)
