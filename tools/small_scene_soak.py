"""One-off soak of the small-scene kernels: random small scenes (tests/test_gpu_parity._random_spec: spheres and
squares, diagonal and general transforms, all shaders, root / orbit camera, ragged sizes, S in {1,2,3,4,5,8}) on
every kernel choice (ray threads, pixel threads, general kernel) against the canonical-order C oracle: hit_index
and tmin bit for bit, pixels rtol 1e-4, fused loss / gradient at the test-suite bars.
usage: python tools/small_scene_soak.py [first_seed] [count]"""
import os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, 'tests'))
from dataclasses import replace
import numpy as np, torch
from oracle import oracle_c as oc
from reversible_raytracer_b200 import render as R
import helpers
from helpers import to_device
import test_gpu_parity as TP
first, count = (int(sys.argv[1]) if len(sys.argv) > 1 else 50000), (int(sys.argv[2]) if len(sys.argv) > 2 else 300)
dev = torch.device('cuda')
bad, runs = 0, 0
for seed in range(first, first + count):
    rng = np.random.RandomState(seed)
    ps = oc.PackedScene.from_spec(TP._random_spec(rng), camera_grad=1)
    img_o, hit_o, tmin_o = oc.render_forward(ps)
    target = np.clip(img_o * 0.6 + 0.2, 0, 1).astype(np.float32)
    image_o, _, loss_o, grad_o = oc.render_fused_mse(ps, target)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, dev)
    for name, c in (('ray', replace(cfg, pixel_threads=2)), ('pixel', replace(cfg, pixel_threads=1)), ('general', replace(cfg, no_small=1))):
        img, hit, tmin = R.render_forward(c, ot, w2o, mat, light, cam, jit, want_hit=True, want_tmin=True)
        loss, grad, image, _ = R.render_fused_mse(c, ot, w2o, mat, light, cam, torch.from_numpy(target).to(dev), None, jit, want_image=True)
        ok = np.array_equal(hit.cpu().numpy().reshape(hit_o.shape), hit_o) and \
            np.array_equal(tmin.cpu().numpy().reshape(tmin_o.shape).view(np.int32), tmin_o.view(np.int32)) and \
            np.allclose(img.cpu().numpy().reshape(img_o.shape), img_o, rtol=1e-4, atol=1e-5) and \
            np.allclose(image.cpu().numpy().reshape(image_o.shape), image_o, rtol=1e-4, atol=1e-5) and \
            abs(float(loss.sum()) - float(loss_o.sum())) <= 1e-4 * max(1.0, abs(float(loss_o.sum())))
        g, r = grad.cpu().numpy().astype(np.float64).reshape(-1), grad_o.reshape(-1)
        sc = np.max(np.abs(r))
        ok = ok and (sc == 0 or np.max(np.abs(g - r)) <= 1e-3 * sc)
        runs += 1
        if not ok:
            bad += 1
            print('MISMATCH seed', seed, name, 'n', ps.n, 'S', ps.samples, 'N', ps.N)
print('small-scene soak: %d scenes x 3 kernel choices (seeds %d..%d): %d of %d runs off the bars' % (count, first, first + count - 1, bad, runs))
