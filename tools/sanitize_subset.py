"""A small end-to-end exercise of every kernel / path for compute-sanitizer runs
(memcheck, racecheck, synccheck): both render kernels in all three modes, chunked tables,
TMA-staged records, culling, shadows, batched scenes, chain kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from dataclasses import replace
from oracle import oracle_c as oc, scenes
from reversible_raytracer_b200 import render as R
from helpers import to_device
dev = torch.device('cuda')
specs = {
    'c3_small': scenes.match_mirror(n=48),
    'orbit': scenes.orbit((3.8, -8.1, 32), 0, n=32),
    'chunked_mixed': None,
    'shadows': scenes.shadow_scene(n=32),
    'many_shadows': None,
}
sp = scenes.stress(n=16, num_objects=600, general=True); sp['obj_type'] = sp['obj_type'].copy(); sp['obj_type'][::7] = 1
specs['chunked_mixed'] = sp
sp = scenes.stress(n=16, num_objects=600); sp['shadows'] = 1
specs['many_shadows'] = sp
for name, spec in specs.items():
    ps = oc.PackedScene.from_spec(spec, camera_grad=1)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, dev)
    for variant in (dict(), dict(no_small=1), dict(no_small=1, use_records=0), dict(no_small=1, cull=1)):
        c = replace(cfg, **variant)
        img, hit, tmin = R.render_forward(c, ot, w2o, mat, light, cam, jit, want_hit=True, want_tmin=True)
        if c.samples <= 8:
            R.render_fused_mse(c, ot, w2o, mat, light, cam, torch.zeros_like(img), None, jit, want_image=True, want_hit=True)
        dl = torch.randn_like(img)
        R.render_backward(c, ot, w2o, mat, light, cam, dl, hit, jit)
        R.render_backward(c, ot, w2o, mat, light, cam, dl, None, jit)
    torch.cuda.synchronize()
    print('ok', name)
# API path: chain kernels + captured optimiser step
from tools.latency import c3
train, sc = c3(True)
for _ in range(5):
    train()
torch.cuda.synchronize()
print('ok api')
