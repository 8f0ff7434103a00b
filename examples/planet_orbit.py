"""B200 port of orbit_experiments/planet_orbit.py: renders n random planet-orbit
scenes from two cameras into a uint8 dataset (n, 2, 64, 64, 3) -- as ONE batched
kernel launch instead of 2n compiled-function calls."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reversible_raytracer_b200 import render as R, workloads as W  # noqa: E402


def main(n=100, x_dims=64, seed=None, out='orbit_dataset.npz'):
    tb = W.orbit_tables(n, seed=seed if seed is not None else np.random.randint(1 << 30))
    dev = torch.device('cuda')
    t = lambda a: torch.from_numpy(a).to(dev)
    cfg = R.RenderConfig(n=x_dims, samples=4, shader=tb['shader'], transpose=tb['transpose'],
                         seed=int(np.random.randint(1 << 30)))
    image, _, _ = R.render_forward(cfg, t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']),
                                   t(tb['camera']), None, want_hit=False)
    dataset = (image.reshape(n, 2, x_dims, x_dims, 3) * 255).to(torch.uint8).cpu().numpy()   # planet_orbit.py:61
    np.savez(out, dataset)
    np.savez(out.replace('dataset', 'target'), tb['centres'].astype(np.float32))
    return dataset


if __name__ == '__main__':
    main()
