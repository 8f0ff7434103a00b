"""B200 port of the reference's generate_data.py: n depth-map renders (x_dims x y_dims, uint8) of
one sphere at a random centre, saved as dataset.npz -- one BATCHED launch instead of n compiled
calls (generate_data.py:20-52)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reversible_raytracer_b200 import render as R, _native as nat  # noqa: E402


def main(n=100, x_dims=32, seed=None, out='dataset.npz'):
    rng = np.random.RandomState(seed)
    centres = np.stack([rng.rand(n) * 4 - 2, rng.rand(n) * 4 - 2, rng.rand(n) * 2 + 4], 1).astype(np.float32)   # :41-42
    dev = torch.device('cuda')
    c = torch.from_numpy(centres).to(dev)
    w2o = R.w2o_translate_scale(c[:, None, :], torch.ones((n, 1, 3), device=dev))                  # translate(center1)
    mat = torch.tensor([[0.5, 0.7, 0.3, 50., 0.2, 0.9, 0.4]], device=dev)                            # material1 (:26)
    light = torch.tensor([-1., -1., 2., 0.961, 1., 0.87], device=dev)
    cam = torch.tensor(np.concatenate([np.eye(4, dtype=np.float32)[:3].reshape(-1), [0, 0, 1]]).astype(np.float32), device=dev)
    cfg = R.RenderConfig(n=x_dims, samples=4, shader=nat.SHADER_DEPTH, transpose=1, max_depth=6.1,
                         seed=int(rng.randint(1 << 30)))
    image, _, _ = R.render_forward(cfg, torch.zeros(1, dtype=torch.int32, device=dev), w2o, mat, light, cam, None,
                                   want_hit=False)
    dataset = (image[..., 0].clamp(0, 1) * 255).to(torch.uint8).cpu().numpy()                        # :49-50
    np.savez(out, dataset)
    return dataset, centres


if __name__ == '__main__':
    main()
