"""Python-3 / B200 port of the reference's optimize_brightness.py (C1): two spheres, a
directional light, Phong shading, 128x128, 4 AA samples; gradient descent on the two
centres to brighten two marker pixels.  Edits vs the reference script are exactly the
ones listed in INTEGRATION.md section 1."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reversible_raytracer_b200.optimize import GDOptimizer  # noqa: E402
from reversible_raytracer_b200.scene import *  # noqa: E402,F401,F403
from reversible_raytracer_b200.shader import *  # noqa: E402,F401,F403


def main(steps=90, out='output', dump=True):
    os.makedirs(out, exist_ok=True)
    center1 = torch.tensor([-.5, -.5, 4.], device='cuda')
    center2 = torch.tensor([.5, .5, 4.], device='cuda')
    material1 = Material((0.2, 0.9, 0.4), 0.3, 0.7, 0.5, 50.)
    material2 = Material((0.87, 0.1, 0.507), 0.3, 0.9, 0.4, 50.)
    t1 = translate(center1)
    t2 = translate(center2) * rotate(90, (0, 0, 1)) * scale((1, 2, 1.5))
    shapes = [Sphere(t1, material1), Sphere(t2, material2)]
    light = Light((-1., -1., 2.), (0.961, 1., 0.87))
    scene = Scene(shapes, [light], Camera(128, 128), PhongShader())

    print('Rendering initial scene')
    render_fn = lambda: scene.build().detach()
    if dump:
        drawWithMarkers(os.path.join(out, '0.png'), render_fn())

    def loss():
        image = scene.build()
        return -image[90, 85].sum() - image[50, 90].sum()

    train = GDOptimizer().optimize([center1, center2], loss, 0.0008, 0.1)
    losses = []
    for i in range(steps):
        losses.append(train())
        print('Step', i + 1, losses[-1])
        if dump:
            drawWithMarkers(os.path.join(out, '%d.png' % (i + 1,)), render_fn())
    return losses, center1, center2


if __name__ == '__main__':
    main()
