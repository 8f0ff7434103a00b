"""Python-3 / B200 port of the reference's match_mirror.py (C3): two spheres and a
square; the two centres are optimised so that the image matches its own left-right
flip.  `--fused` uses the single-kernel forward+loss+reverse path (Scene.build_mse).
`--mirror` makes the square an actual mirror (Material(..., reflectivity=0.8): one reflection
bounce, RRT_FLAG_MIRROR -- an extension, the reference has no secondary ray) and the spheres
are then also matched through their reflections."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _common as C  # noqa: E402
from reversible_raytracer_b200.optimize import GDOptimizer  # noqa: E402
from reversible_raytracer_b200.scene import *  # noqa: E402,F401,F403
from reversible_raytracer_b200.shader import *  # noqa: E402,F401,F403


def build_scene(params, mirror=False):
    green, pink = C.materials()
    wall_mat = Material((0.87, 0.1, 0.507), 0.3, 0.9, 0.4, 50., reflectivity=0.8) if mirror else pink
    wall = Square(translate((0, 0, 3)) * rotate(50, [0., 1., 0.]), wall_mat)
    balls = [Sphere(translate(p), m) for p, m in zip(params, (green, pink))]
    return Scene(balls + [wall], [Light((-1., -1., 2.), (1., 0.87, 0.961))], Camera(128, 128), PhongShader())


def main(steps=90, out='output', dump=True, fused=False, mirror=False):
    params = C.centres()
    scene = build_scene(params, mirror)
    first = scene.build().detach()
    mirrored = torch.flip(first, dims=[1])                      # np.fliplr, match_mirror.py:40
    if dump:
        os.makedirs(out, exist_ok=True)
        draw(os.path.join(out, '0.png'), first)
        draw(os.path.join(out, '0lr.png'), mirrored)
    # match_mirror.py:45 -- as tensor algebra, or as Scene.mse_cost (whole step = one kernel launch)
    cost = scene.mse_cost(mirrored) if fused else (lambda: ((scene.build() - mirrored) ** 2).sum())
    train = GDOptimizer().optimize(params, cost, 0.000008, 0.1)
    losses = C.run(train, steps, lambda: scene.build().detach(), out if dump else None, draw)
    return losses, params[0], params[1]


if __name__ == '__main__':
    main(fused='--fused' in sys.argv, mirror='--mirror' in sys.argv)
