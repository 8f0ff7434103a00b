"""Python-3 / B200 port of the reference's test_balls.py (BASELINE config C2): an MLP encoder
(1024 -> 300 -> 30 -> 10 -> 2 capsules x 6) is trained so that the DEPTH-MAP render of two
spheres `translate(p[:3]) * scale(p[3:])` reproduces one 32x32 target image.  Structure and
hyper-parameters follow test_balls.py:17-76; the edits are the ones INTEGRATION.md lists
(eager tensors instead of theano.shared, the golden 15.jpg from tests/golden since the
script's 15.png does not exist in the reference tree)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from autoencoder import Autoencoder  # noqa: E402
from reversible_raytracer_b200.scene import *  # noqa: E402,F401,F403
from reversible_raytracer_b200.shader import *  # noqa: E402,F401,F403
from reversible_raytracer_b200.optimize import MGDAutoOptimizer  # noqa: E402
from reversible_raytracer_b200.util import get_epsilon, draw  # noqa: E402


def main(num_epoch=200, epsilon=0.0001, num_capsule=2, out_dir=None, verbose=True, graph='auto'):
    dev = torch.device('cuda')
    img = np.load(os.path.join(ROOT, 'tests', 'golden', 'balls_15.npy')).astype(np.float32)
    if img.ndim == 3:
        img = img[:, :, 0]
    train_data = torch.tensor(img.reshape(1, -1) / 255.0, device=dev)              # test_balls.py:17
    N, D = train_data.shape
    img_sz = int(np.sqrt(D))

    def scene(capsules, obj_params):                                                # test_balls.py:22-44
        material1 = Material((0.2, 0.9, 0.4), 0.3, 0.7, 0.5, 50.)
        shapes = []
        for capsule, obj_param in zip(capsules, obj_params):
            t1 = translate(obj_param[:3]) * scale(obj_param[3:])
            shapes.append(Sphere(t1, material1) if capsule.name == 'sphere' else Square(t1, material1))
        light = Light((-1., -1., 2.), (0.961, 1., 0.87))
        return Scene(shapes, [light], Camera(img_sz, img_sz), DepthMapShader(6.1)).build(seed=15)

    ae = Autoencoder(scene, D, 300, 30, 10, num_capsule, device=dev)
    train_ae = MGDAutoOptimizer(ae).optimize(train_data, graph=graph)     # one CUDA-graph replay per epoch after two eager ones
    main.last_state = train_ae.state
    losses = []
    for n in range(1, num_epoch + 1):
        eps = get_epsilon(epsilon, num_epoch, n)
        losses.append(train_ae(eps))
        if verbose and (n % 20 == 0 or n == 1):
            with torch.no_grad():
                c = ae.encoder(train_data[0])[0]
            print('...Epoch %d Train loss %g, Center (%g, %g, %g)' % (n, losses[-1], c[0], c[1], c[2]))
        if out_dir and n % 10 == 0:
            with torch.no_grad():
                draw(os.path.join(out_dir, 'test_balls%d.png' % n), ae.get_reconstruct(train_data[0])[:, :, 0])
    return losses


if __name__ == '__main__':
    main()
