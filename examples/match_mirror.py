"""Python-3 / B200 port of the reference's match_mirror.py (C3): two spheres and a
square; the two centres are optimised so that the image matches its own left-right
flip.  `--fused` uses the single-kernel forward+loss+reverse path (Scene.build_mse)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reversible_raytracer_b200.optimize import GDOptimizer  # noqa: E402
from reversible_raytracer_b200.scene import *  # noqa: E402,F401,F403
from reversible_raytracer_b200.shader import *  # noqa: E402,F401,F403


def main(steps=90, out='output', dump=True, fused=False):
    os.makedirs(out, exist_ok=True)
    center1 = torch.tensor([-.5, -.5, 4.], device='cuda')
    center2 = torch.tensor([.5, .5, 4.], device='cuda')
    material1 = Material((0.2, 0.9, 0.4), 0.3, 0.7, 0.5, 50.)
    material2 = Material((0.87, 0.1, 0.507), 0.3, 0.9, 0.4, 50.)
    objs = [
        Sphere(translate(center1), material1),
        Sphere(translate(center2), material2),
        Square(translate((0, 0, 3)) * rotate(50, [0., 1., 0.]), material2),
    ]
    light = Light((-1., -1., 2.), (1., 0.87, 0.961))
    scene = Scene(objs, [light], Camera(128, 128), PhongShader())

    print('Rendering initial scene')
    render = scene.build().detach()
    flipped = torch.flip(render, dims=[1])                      # np.fliplr, match_mirror.py:40
    if dump:
        draw(os.path.join(out, '0.png'), render)
        draw(os.path.join(out, '0lr.png'), flipped)

    if fused:
        cost = lambda: scene.build_mse(flipped)
    else:
        cost = lambda: ((scene.build() - flipped) ** 2).sum()  # match_mirror.py:45
    train = GDOptimizer().optimize([center1, center2], cost, 0.000008, 0.1)
    losses = []
    for i in range(steps):
        losses.append(train())
        print('Step', i + 1, losses[-1])
        if dump:
            draw(os.path.join(out, '%d.png' % (i + 1,)), scene.build().detach())
    return losses, center1, center2


if __name__ == '__main__':
    main(fused='--fused' in sys.argv)
