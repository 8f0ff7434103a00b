import numpy as np, torch, sys
from reversible_raytracer_b200.scene import *
from reversible_raytracer_b200.shader import *
from reversible_raytracer_b200.optimize import GDOptimizer
dev='cuda'
center1 = torch.tensor([-.5,-.5,4.], device=dev, requires_grad=True)
center2 = torch.tensor([.5,.5,4.], device=dev, requires_grad=True)
material1 = Material((0.2, 0.9, 0.4), 0.3, 0.7, 0.5, 50.)
material2 = Material((0.87, 0.1, 0.507), 0.3, 0.9, 0.4, 50.)
t1 = lambda: translate(center1)
shapes = lambda: [Sphere(translate(center1), material1), Sphere(translate(center2) * rotate(90, (0, 0, 1)) * scale((1, 2, 1.5)), material2)]
light = Light((-1., -1., 2.), (0.961, 1., 0.87))
camera = Camera(128, 128)
scene = Scene(shapes(), [light], camera, PhongShader())
def loss():
    scene.shapes = shapes()
    image = scene.build()
    return -image[90, 85].sum() - image[50, 90].sum()
train = GDOptimizer().optimize([center1, center2], loss, 0.0008, 0.1)
for i in range(5): print(i, train(), center1.tolist())
