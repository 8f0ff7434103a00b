"""Python-3 / B200 port of the reference's optimize_brightness.py (C1): two spheres, a
directional light, Phong shading, 128x128, 4 AA samples; gradient descent on the two
centres to brighten two marker pixels.  Edits vs the reference script are exactly the
ones listed in INTEGRATION.md section 1."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402
import _common as C  # noqa: E402
from reversible_raytracer_b200.optimize import GDOptimizer  # noqa: E402
from reversible_raytracer_b200.scene import *  # noqa: E402,F401,F403
from reversible_raytracer_b200.shader import *  # noqa: E402,F401,F403

MARKERS = ((90, 85), (50, 90))                                   # optimize_brightness.py:51


def build_scene(params):
    green, pink = C.materials()
    ellipsoid = translate(params[1]) * rotate(90, (0, 0, 1)) * scale((1, 2, 1.5))
    return Scene([Sphere(translate(params[0]), green), Sphere(ellipsoid, pink)],
                 [Light((-1., -1., 2.), (0.961, 1., 0.87))], Camera(128, 128), PhongShader())


def main(steps=90, out='output', dump=True, fused=False):
    """fused=True: the same loss written as a weight image (-1 at the two markers) through
    Scene.linear_cost -- GDOptimizer then runs the whole step as ONE kernel launch."""
    params = C.centres()
    scene = build_scene(params)
    frame = lambda: scene.build().detach()
    if dump:
        os.makedirs(out, exist_ok=True)
        drawWithMarkers(os.path.join(out, '0.png'), frame())

    def brightness():                                            # minimised: minus the markers' brightness
        image = scene.build()
        return -sum(image[a, b].sum() for a, b in MARKERS)

    if fused:
        weights = torch.zeros((128, 128, 3), device=params[0].device)
        for a, b in MARKERS:
            weights[a, b] = -1.0
        brightness = scene.linear_cost(weights)
    train = GDOptimizer().optimize(params, brightness, 0.0008, 0.1)
    losses = C.run(train, steps, frame, out if dump else None, drawWithMarkers)
    return losses, params[0], params[1]


if __name__ == '__main__':
    main(fused='--fused' in sys.argv)
