"""GPU parity: CUDA kernels (through the C ABI) vs the canonical-order C oracle.

Bars (BASELINE.json north_star): hit/object-index masks BIT-EXACT; pixels within
rtol 1e-4 (+ atol 1e-5, images live in [0,1]); gradients within 1e-3 of each
parameter block's max magnitude.
"""
from dataclasses import replace

import numpy as np
import pytest
import torch

from oracle import oracle_c as oc, scenes
from reversible_raytracer_b200 import render as R, workloads as W
import helpers
from helpers import to_device, block_rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(params=['auto', 'general', 'norecords', 'pixel', 'ray'], autouse=True)
def kernel_choice(request, monkeypatch):
    """Every case runs five times: with the library's own choices (small scenes take the small-scene
    kernel -- one ray per thread for a single image, one pixel per thread for big batches --, >= 64
    objects get a prebuilt record table staged by TMA), with the general kernel forced
    (RRT_FLAG_NO_SMALL), with the record table switched off (the kernels build the sweep records
    per CTA), and with each thread mapping of the small-scene kernel forced (RRT_FLAG_PIXEL_THREADS
    / RRT_FLAG_RAY_THREADS)."""
    monkeypatch.setattr(helpers, 'NO_SMALL', request.param == 'general')
    monkeypatch.setattr(helpers, 'USE_RECORDS', request.param != 'norecords')
    monkeypatch.setattr(helpers, 'PIXEL_THREADS', {'pixel': 1, 'ray': 2}.get(request.param, 0))
    return request.param

PIX_RTOL, PIX_ATOL, GRAD_TOL = 1e-4, 1e-5, 1e-3

CASES = {
    'C1_optimize_brightness': lambda: scenes.optimize_brightness(),
    'C2_test_balls_depth': lambda: scenes.test_balls(),
    'C3_match_mirror_square': lambda: scenes.match_mirror(),
    'C4_orbit_view0': lambda: scenes.orbit((3.8307, -8.1441, 32), 0),
    'C4_orbit_view1': lambda: scenes.orbit((-5.0, 7.4833, 32), 1),
    'C5_stress_diag': lambda: scenes.stress(n=128, num_objects=100),
    'C5g_stress_general': lambda: scenes.stress(n=96, num_objects=48, general=True),
    'S1': lambda: scenes.stress(n=100, num_objects=20, samples=1),
    'S2': lambda: scenes.stress(n=70, num_objects=20, samples=2),
    'S8': lambda: scenes.stress(n=40, num_objects=20, samples=8),
    'S3_generic': lambda: scenes.stress(n=33, num_objects=9, samples=3),
    'S16_generic': lambda: scenes.stress(n=20, num_objects=9, samples=16),
    'ragged_n1': lambda: scenes.optimize_brightness(n=1),
    'ragged_n5': lambda: scenes.match_mirror(n=5),
    'many_objects_chunked': lambda: scenes.stress(n=32, num_objects=1500),
    'many_general_chunked': lambda: scenes.stress(n=24, num_objects=1100, general=True),
    'many_mixed_chunked': lambda: _mixed_many(),
}


def _mixed_many():
    """> 512 objects (two table chunks) with squares and general spheres interleaved."""
    from oracle import oracle_numpy as on
    base = scenes.stress(n=20, num_objects=600, general=True)
    base['obj_type'] = base['obj_type'].copy()
    base['obj_type'][::7] = on.SQUARE
    return base


def check_forward(ps, dev):
    img_o, hit_o, tmin_o = oc.render_forward(ps)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, dev)
    img, hit, tmin = R.render_forward(cfg, ot, w2o, mat, light, cam, jit, want_hit=True, want_tmin=True)
    torch.cuda.synchronize()
    img, hit, tmin = img.cpu().numpy(), hit.cpu().numpy(), tmin.cpu().numpy()
    assert np.array_equal(hit.reshape(hit_o.shape), hit_o), 'hit masks must be bit-exact'
    assert np.array_equal(tmin.reshape(tmin_o.shape).view(np.uint32), tmin_o.view(np.uint32)), 'tmin must be bit-exact'
    np.testing.assert_allclose(img.reshape(img_o.shape), img_o, rtol=PIX_RTOL, atol=PIX_ATOL)
    return img_o, hit_o


@pytest.mark.parametrize('name', list(CASES))
def test_forward_parity(name, cuda):
    ps = oc.PackedScene.from_spec(CASES[name](), camera_grad=1)
    check_forward(ps, cuda)


def compare_grads(flat, ref, N):
    g, r = oc.split_grad(flat, N), oc.split_grad(ref, N)
    for k in range(N):
        assert block_rel_err(g['w2o'][k], r['w2o'][k]) <= GRAD_TOL or np.max(np.abs(r['w2o'])) * 1e-5 > np.max(np.abs(r['w2o'][k])), ('w2o', k)
    assert block_rel_err(g['w2o'], r['w2o']) <= GRAD_TOL
    assert block_rel_err(g['material'], r['material']) <= GRAD_TOL
    for key in ('light_dir', 'light_int', 'cam_o2w', 'look_at'):
        assert block_rel_err(g[key], r[key]) <= GRAD_TOL, key


@pytest.mark.parametrize('name', [c for c in CASES if 'generic' not in c])
def test_fused_mse_parity(name, cuda):
    ps = oc.PackedScene.from_spec(CASES[name](), camera_grad=1)
    img_o, _, _ = oc.render_forward(ps, want_aux=False)
    rng = np.random.RandomState(11)
    target = np.clip(img_o + rng.normal(0, 0.1, img_o.shape), 0, 1).astype(np.float32)
    cw = (1.0, 0.0, 0.0) if 'depth' in name else None
    image_o, hit_o, loss_o, grad_o = oc.render_fused_mse(ps, target, cw)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, cuda)
    loss, grad, image, hit = R.render_fused_mse(cfg, ot, w2o, mat, light, cam, torch.from_numpy(target).to(cuda),
                                                cw, jit, want_image=True, want_hit=True)
    torch.cuda.synchronize()
    assert np.array_equal(hit.cpu().numpy().reshape(hit_o.shape), hit_o)
    np.testing.assert_allclose(image.cpu().numpy().reshape(image_o.shape), image_o, rtol=PIX_RTOL, atol=PIX_ATOL)
    np.testing.assert_allclose(float(loss), loss_o[0], rtol=1e-4)
    compare_grads(grad.cpu().numpy().astype(np.float64), grad_o[0], ps.N)


@pytest.mark.parametrize('name', list(CASES))
@pytest.mark.parametrize('stored', [True, False])
def test_backward_parity(name, stored, cuda):
    ps = oc.PackedScene.from_spec(CASES[name](), camera_grad=1)
    img_o, hit_o, _ = oc.render_forward(ps)
    rng = np.random.RandomState(5)
    dl = rng.normal(0, 1, img_o.shape).astype(np.float32)
    grad_o = oc.render_backward(ps, dl, hit_o)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, cuda)
    hit = torch.from_numpy(hit_o).to(cuda) if stored else None
    grad = R.render_backward(cfg, ot, w2o, mat, light, cam, torch.from_numpy(dl).to(cuda), hit, jit)
    torch.cuda.synchronize()
    compare_grads(grad.cpu().numpy().astype(np.float64), grad_o[0], ps.N)


def test_batched_scenes(cuda):
    """C4 shape: a batch of scenes x 2 camera views, per-scene tables."""
    rng = np.random.RandomState(1234)
    specs = []
    for q in range(6):
        th = rng.uniform(0, 2 * np.pi)
        specs.append(scenes.orbit((9 * np.cos(th), 9 * np.sin(th), 32), q % 2, seed=100 + q))
    pss = [oc.PackedScene.from_spec(s, camera_grad=1) for s in specs]
    B = len(pss)
    ps = oc.PackedScene(64, 4, pss[0].obj_type, np.stack([p.w2o[0] for p in pss]),
                        np.stack([p.material[0] for p in pss]), np.stack([p.light[0] for p in pss]),
                        np.stack([p.camera[0] for p in pss]), pss[0].shader, 0,
                        jitter_x=np.stack([p.jitter_x for p in pss]), jitter_y=np.stack([p.jitter_y for p in pss]),
                        camera_grad=1)
    img_o, hit_o, _ = oc.render_forward(ps)
    target = np.clip(img_o + rng.normal(0, 0.1, img_o.shape), 0, 1).astype(np.float32)
    image_o, hit_o, loss_o, grad_o = oc.render_fused_mse(ps, target)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, cuda)
    loss, grad, image, hit = R.render_fused_mse(cfg, ot, w2o, mat, light, cam, torch.from_numpy(target).to(cuda),
                                                None, jit, want_image=True, want_hit=True)
    assert np.array_equal(hit.cpu().numpy(), hit_o)
    np.testing.assert_allclose(image.cpu().numpy(), image_o, rtol=PIX_RTOL, atol=PIX_ATOL)
    np.testing.assert_allclose(loss.cpu().numpy(), loss_o, rtol=1e-4)
    for b in range(B):
        compare_grads(grad[b].cpu().numpy().astype(np.float64), grad_o[b], ps.N)


def test_row_slabs_and_rng_jitter(cuda):
    """Multi-GPU sharding unit: row slabs reproduce the full image bit-exactly, with
    the in-kernel counter RNG (no jitter buffers); slab gradients sum to the full one."""
    spec = scenes.stress(n=96, num_objects=40)
    full = oc.PackedScene.from_spec(spec, camera_grad=0, use_rng_seed=777)
    img_o, hit_o, _ = oc.render_forward(full)
    cfg, ot, w2o, mat, light, cam, _ = to_device(full, cuda, with_jitter=False)
    img, hit, _ = R.render_forward(cfg, ot, w2o, mat, light, cam, None)
    assert np.array_equal(hit.cpu().numpy().reshape(hit_o.shape), hit_o)
    np.testing.assert_allclose(img.cpu().numpy().reshape(img_o.shape), img_o, rtol=PIX_RTOL, atol=PIX_ATOL)
    target = torch.zeros_like(img)
    loss_f, grad_f, _, _ = R.render_fused_mse(cfg, ot, w2o, mat, light, cam, target)
    parts, gsum, lsum = [], 0, 0
    for rb, rc in ((0, 40), (40, 13), (53, 43)):
        c2 = cfg.slab(rb, rc)
        im, _, _ = R.render_forward(c2, ot, w2o, mat, light, cam, None)
        parts.append(im)
        l, g, _, _ = R.render_fused_mse(c2, ot, w2o, mat, light, cam, target[rb:rb + rc])
        gsum, lsum = gsum + g.double(), lsum + l
    assert torch.equal(torch.cat(parts, 0), img)
    np.testing.assert_allclose(float(lsum), float(loss_f), rtol=1e-5)
    N = full.N
    compare_grads(gsum.cpu().numpy(), grad_f.double().cpu().numpy(), N)


def test_error_reporting(cuda):
    from reversible_raytracer_b200 import _native as nat
    ps = oc.PackedScene.from_spec(scenes.test_balls(), camera_grad=0)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, cuda)
    bad = R.RenderConfig(n=32, samples=0)
    with pytest.raises(nat.NativeError):
        R.render_forward(bad, ot, w2o, mat, light, cam, None)
    with pytest.raises(nat.NativeError):
        R.render_forward(cfg, ot.cpu(), w2o.cpu(), mat.cpu(), light.cpu(), cam.cpu(), None)


def test_ray_table_and_in_kernel_grid_agree_bitwise(cuda, monkeypatch):
    """rrt_scene.base_rays (precomputed by rrt_primary_rays, used for n <= 1024) and the
    in-kernel float64 grid give identical bits; both equal the oracle's rays."""
    ps = oc.PackedScene.from_spec(scenes.stress(n=96, num_objects=30), camera_grad=0)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, cuda)
    a = R.render_forward(cfg, ot, w2o, mat, light, cam, jit, want_hit=True, want_tmin=True)
    monkeypatch.setattr(R, 'BASE_RAYS_MAX_N', 0)
    b = R.render_forward(cfg, ot, w2o, mat, light, cam, jit, want_hit=True, want_tmin=True)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    tab = R.base_rays(96, cuda).cpu().numpy()
    from oracle import oracle_numpy as on
    _, ref = on.make_rays(96, 96)
    assert np.array_equal(tab.view(np.uint32), ref.view(np.uint32))


def test_full_resolution_row_slab(cuda):
    """BASELINE's full image side (n = 4096, in-kernel ray grid, counter RNG) on a thin
    row slab: masks bit-exact, pixels and gradients within tolerance."""
    spec = scenes.stress(n=8, num_objects=64)        # tables only; n is overridden below
    ps0 = oc.PackedScene.from_spec(spec, camera_grad=0, use_rng_seed=4321)
    ps = oc.PackedScene(4096, 4, ps0.obj_type, ps0.w2o, ps0.material, ps0.light, ps0.camera, ps0.shader, 1,
                        seed=4321, row_begin=2045, row_count=6)
    img_o, hit_o, tmin_o = oc.render_forward(ps)
    cfg, ot, w2o, mat, light, cam, _ = to_device(ps, cuda, with_jitter=False)
    img, hit, tmin = R.render_forward(cfg, ot, w2o, mat, light, cam, None, want_hit=True, want_tmin=True)
    assert np.array_equal(hit.cpu().numpy().reshape(hit_o.shape), hit_o)
    assert np.array_equal(tmin.cpu().numpy().reshape(tmin_o.shape).view(np.uint32), tmin_o.view(np.uint32))
    np.testing.assert_allclose(img.cpu().numpy().reshape(img_o.shape), img_o, rtol=PIX_RTOL, atol=PIX_ATOL)
    target = np.zeros_like(img_o)
    _, _, loss_o, grad_o = oc.render_fused_mse(ps, target)
    loss, grad, _, _ = R.render_fused_mse(cfg, ot, w2o, mat, light, cam, torch.from_numpy(target[0]).to(cuda))
    np.testing.assert_allclose(float(loss), loss_o[0], rtol=1e-4)
    compare_grads(grad.cpu().numpy().astype(np.float64), grad_o[0], ps.N)


def test_empty_scene(cuda):
    cfg = R.RenderConfig(n=16, samples=4)
    z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=cuda)
    light = torch.tensor([-1., -1., 2., 1., 1., 1.], device=cuda)
    cam = torch.tensor([1., 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 1], device=cuda)
    img, hit, _ = R.render_forward(cfg, torch.zeros(0, dtype=torch.int32, device=cuda), z(0, 12), z(0, 7), light, cam, None)
    assert float(img.abs().max()) == 0.0 and int((hit >= 0).sum()) == 0
    loss, grad, _, _ = R.render_fused_mse(cfg, torch.zeros(0, dtype=torch.int32, device=cuda), z(0, 12), z(0, 7), light, cam,
                                          torch.ones(16, 16, 3, device=cuda))
    assert abs(float(loss) - 16 * 16 * 3) < 1e-3 and float(grad.abs().max()) == 0.0


def _random_spec(rng):
    """Random small scene: spheres and squares, diagonal and general transforms, any shader,
    root or orbit camera, ragged image sizes, S in {1,2,3,4,5,8}."""
    from oracle import oracle_numpy as on
    n = int(rng.choice([3, 8, 17, 32, 45, 64]))
    S = int(rng.choice([1, 2, 3, 4, 5, 8]))
    N = int(rng.randint(1, 12))
    shapes = []
    for _ in range(N):
        c = (rng.uniform(-1.5, 1.5), rng.uniform(-1.5, 1.5), rng.uniform(2.5, 7))
        t = on.translate(c)
        kind = on.SQUARE if rng.rand() < 0.25 else on.SPHERE
        if rng.rand() < 0.5:
            ax = rng.normal(size=3)
            t = on.compose(t, on.rotate(rng.uniform(0, 360), ax / np.linalg.norm(ax)))
        t = on.compose(t, on.scale(rng.uniform(0.3, 1.2, 3) if rng.rand() < 0.5 else [rng.uniform(0.3, 1.2)] * 3))
        mat = np.array([rng.uniform(.1, .5), rng.uniform(.3, .9), rng.uniform(0, .5), float(rng.choice([1, 8, 50])),
                        *rng.uniform(0.05, 1, 3)], dtype=np.float32)
        shapes.append((kind, t, mat))
    shader = str(rng.choice(['phong', 'phong_nospec', 'depth']))
    cam = None
    if rng.rand() < 0.5:
        cam = on.compose(on.translate(rng.uniform(-0.5, 0.5, 3)), on.rotate(rng.uniform(-10, 10), (0, 1, 0)))
    return scenes.spec_from(n, S, shapes, (rng.normal(size=3) + np.array([0, 0, 2.0]), rng.uniform(0.5, 1, 3)), shader,
                            cam=cam, look_at=(0, 0, 1.), max_depth=float(rng.uniform(4, 9)), seed=int(rng.randint(1 << 30)))


@pytest.mark.parametrize('seed', range(24))
def test_random_scenes_forward_and_gradients(seed, cuda):
    rng = np.random.RandomState(1000 + seed)
    ps = oc.PackedScene.from_spec(_random_spec(rng), camera_grad=1)
    img_o, hit_o = check_forward(ps, cuda)
    dl = rng.normal(0, 1, img_o.shape).astype(np.float32)
    grad_o = oc.render_backward(ps, dl, hit_o)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, cuda)
    grad = R.render_backward(cfg, ot, w2o, mat, light, cam, torch.from_numpy(dl).to(cuda), None, jit)
    g, r = grad.cpu().numpy().astype(np.float64), grad_o[0]
    scale = np.max(np.abs(r))
    if scale > 0:
        # random scenes contain grazing hits whose single-ray gradient dominates a block: compare
        # against the whole vector's scale here; per-block comparisons are in the config tests
        assert np.max(np.abs(g - r)) <= GRAD_TOL * scale, np.max(np.abs(g - r)) / scale


def test_scene_batch_shards_reproduce_the_whole_batch(cuda):
    """Scene-batch sharding (C4 across GPUs): rendering scenes [4,10) as their own call with
    scene_begin=4 gives the same bits as the whole batch (in-kernel jitter is keyed by the
    global scene index); checked against the oracle too."""
    from reversible_raytracer_b200 import workloads as W
    tb = W.orbit_tables(5)                                      # 10 scene-views
    ps = oc.PackedScene(32, 4, tb['obj_type'], tb['w2o'], tb['material'], tb['light'], tb['camera'], tb['shader'], 0,
                        seed=99, camera_grad=1)
    img_o, hit_o, _ = oc.render_forward(ps)
    cfg, ot, w2o, mat, light, cam, _ = to_device(ps, cuda, with_jitter=False)
    img, hit, _ = R.render_forward(cfg, ot, w2o, mat, light, cam, None)
    assert np.array_equal(hit.cpu().numpy(), hit_o)
    from dataclasses import replace
    part, hit_p, _ = R.render_forward(replace(cfg, scene_begin=4), ot, w2o[4:], mat, light, cam[4:], None)
    assert torch.equal(part, img[4:]) and torch.equal(hit_p, hit[4:])
    ps_part = oc.PackedScene(32, 4, tb['obj_type'], tb['w2o'][4:], tb['material'], tb['light'], tb['camera'][4:], tb['shader'], 0,
                             seed=99, camera_grad=1, scene_begin=4)
    _, hit_op, _ = oc.render_forward(ps_part)
    assert np.array_equal(hit_op, hit_o[4:])


def test_full_size_c5_properties(cuda):
    """BASELINE's full stress config (4096 x 4096, S=4, 1024 spheres, in-kernel RNG) through
    size-independent properties: (1) 8 row slabs reproduce the full render bit for bit,
    (2) sampled rows match the oracle bit-exactly (masks) / within tolerance (pixels),
    (3) fused forward+loss+reverse == forward, torch loss, separate reverse pass,
    (4) the reverse pass is linear in dL/dimage, (5) two runs give identical masks."""
    from reversible_raytracer_b200 import workloads as W, _native as nat
    n, S, N = 4096, 4, 1024
    tb = W.stress_tables(N)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    args = (t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']), t(tb['camera']))
    cfg = R.RenderConfig(n=n, samples=S, shader=nat.SHADER_PHONG, transpose=1, seed=4321)
    img, hit, _ = R.render_forward(cfg, *args, None, want_hit=True)
    # (5)
    img2, hit2, _ = R.render_forward(cfg, *args, None, want_hit=True)
    assert torch.equal(hit, hit2) and torch.equal(img, img2)
    del img2, hit2
    # (1)
    for r in range(8):
        part, hpart, _ = R.render_forward(cfg.slab(512 * r, 512), *args, None, want_hit=True)
        assert torch.equal(part, img[512 * r:512 * (r + 1)])
        assert torch.equal(hpart, hit[:, 512 * r:512 * (r + 1)])
        del part, hpart
    # (2)
    for row in (0, 1717, 4095):
        ps = oc.PackedScene(n, S, tb['obj_type'], tb['w2o'], tb['material'], tb['light'], tb['camera'], tb['shader'], 1,
                            seed=4321, row_begin=row, row_count=1)
        img_o, hit_o, _ = oc.render_forward(ps)
        assert np.array_equal(hit[:, row:row + 1].cpu().numpy(), hit_o[0])
        np.testing.assert_allclose(img[row:row + 1].cpu().numpy(), img_o[0], rtol=PIX_RTOL, atol=PIX_ATOL)
    # (3) on a 256-row slab (keeps the test light)
    sl = cfg.slab(1024, 256)
    target = torch.rand(256, n, 3, device=cuda)
    loss_f, grad_f, image_f, _ = R.render_fused_mse(sl, *args, target, want_image=True)
    assert torch.equal(image_f, img[1024:1280])
    dl = 2 * (image_f - target)
    grad_b = R.render_backward(sl, *args, dl, hit[:, 1024:1280].contiguous())
    np.testing.assert_allclose(float(loss_f), float(((image_f.double() - target.double()) ** 2).sum()), rtol=1e-5)
    compare_grads(grad_f.double().cpu().numpy(), grad_b.double().cpu().numpy(), N)
    # (4)
    g1, g2 = torch.randn_like(dl), torch.randn_like(dl)
    h = hit[:, 1024:1280].contiguous()
    ga = R.render_backward(sl, *args, g1, h).double()
    gb = R.render_backward(sl, *args, g2, h).double()
    gc = R.render_backward(sl, *args, 0.5 * g1 - 2.0 * g2, h).double()
    ref = (0.5 * ga - 2.0 * gb).cpu().numpy()
    assert np.max(np.abs(gc.cpu().numpy() - ref)) <= 2e-3 * np.max(np.abs(ref))


def _cull_equal(ps, cuda):
    from dataclasses import replace
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, cuda)
    a = R.render_forward(cfg, ot, w2o, mat, light, cam, jit, want_hit=True, want_tmin=True)
    b = R.render_forward(replace(cfg, cull=1), ot, w2o, mat, light, cam, jit, want_hit=True, want_tmin=True)
    for x, y in zip(a, b):
        assert torch.equal(x, y), 'culling changed the result'
    if cfg.samples <= 8:
        target = torch.zeros_like(a[0])
        l0, g0, _, _ = R.render_fused_mse(cfg, ot, w2o, mat, light, cam, target, None, jit)
        l1, g1, _, _ = R.render_fused_mse(replace(cfg, cull=1), ot, w2o, mat, light, cam, target, None, jit)
        np.testing.assert_allclose(float(l1.sum()), float(l0.sum()), rtol=1e-6)
        s = float(g0.abs().max())
        if s > 0:
            assert float((g1 - g0).abs().max()) <= 1e-4 * s        # only the atomic order differs


@pytest.mark.parametrize('name', list(CASES))
def test_culling_is_bit_identical_configs(name, cuda):
    """RRT_FLAG_CULL (conservative per-tile object culling) never changes a bit of the image,
    the hit masks or tmin."""
    _cull_equal(oc.PackedScene.from_spec(CASES[name](), camera_grad=1), cuda)


@pytest.mark.parametrize('seed', range(16))
def test_culling_is_bit_identical_random(seed, cuda):
    _cull_equal(oc.PackedScene.from_spec(_random_spec(np.random.RandomState(5000 + seed)), camera_grad=1), cuda)


def test_culling_full_size_slab(cuda):
    """C5 at full resolution (rows where culling rejects almost everything) + extreme
    geometry: spheres behind the camera (negative t still hits), huge and tiny spheres."""
    from dataclasses import replace
    from reversible_raytracer_b200 import workloads as W, _native as nat
    tb = W.stress_tables(1024)
    w = tb['w2o'].reshape(-1, 3, 4).copy()
    w[5] = W.w2o_translate_scale(np.array([[0.3, -0.2, -6.0]]), np.array([[1.5, 1.5, 1.5]])).reshape(3, 4)     # behind the camera
    w[6] = W.w2o_translate_scale(np.array([[0.0, 0.0, 40.0]]), np.array([[30., 30., 30.]])).reshape(3, 4)      # huge, far
    w[7] = W.w2o_translate_scale(np.array([[0.01, 0.02, 3.0]]), np.array([[1e-3, 1e-3, 1e-3]])).reshape(3, 4)  # sub-pixel
    w[8] = W.w2o_translate_scale(np.array([[0.0, 0.0, 0.5]]), np.array([[2., 2., 2.]])).reshape(3, 4)          # camera inside
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    args = (t(tb['obj_type']), t(w.reshape(-1, 12)), t(tb['material']), t(tb['light']), t(tb['camera']))
    for rb in (0, 2040, 4032):
        cfg = R.RenderConfig(n=4096, samples=4, shader=nat.SHADER_PHONG, transpose=1, seed=4321, row_begin=rb, row_count=64)
        a = R.render_forward(cfg, *args, None, want_hit=True, want_tmin=True)
        b = R.render_forward(replace(cfg, cull=1), *args, None, want_hit=True, want_tmin=True)
        for x, y in zip(a, b):
            assert torch.equal(x, y)


@pytest.mark.parametrize('num_objects', [40, 100, 600])
def test_streamed_host_buffers_equal_single_launch(num_objects, cuda):
    """render.StreamedFusedMSE (row slabs pipelined over copy-in / kernel / copy-out streams,
    target and image in pinned HOST memory) == one whole-image launch: image bit-identical,
    loss / gradients up to the summation order.  40 objects: records built per CTA; 100: the
    record table built once per call + TMA staging on the side streams (what bench.py's e2e leg
    runs); 600: two TMA chunks.  The scene CHANGES between calls, so a record table that is stale
    or read before it is rebuilt shows up as a different image."""
    spec = scenes.stress(n=96, num_objects=num_objects, samples=4, seed=11)
    ps = oc.PackedScene.from_spec(spec, camera_grad=0)
    cfg, ot, w2o, mat, light, cam, _ = to_device(ps, cuda, with_jitter=False)
    cfg = R.RenderConfig(n=cfg.n, samples=cfg.samples, shader=cfg.shader, transpose=cfg.transpose, seed=77,
                         no_small=cfg.no_small, use_records=cfg.use_records)
    w2o_b = w2o.clone()
    w2o_b[:, 3] += 0.37 * w2o_b[:, 0]             # every centre moves: another image
    target = torch.rand((cfg.n, cfg.n, 3), device=cuda)
    ref = [R.render_fused_mse(cfg, ot, w, mat, light, cam, target, want_image=True) for w in (w2o, w2o_b)]
    assert not torch.equal(ref[0][2], ref[1][2])
    pin_t = target.cpu().pin_memory()
    pin_i = torch.empty_like(pin_t).pin_memory()
    # uniform slab counts, and an explicit graded schedule (short slabs at both ends, the kind the automatic
    # choice takes for tall images) including a slab that is not a multiple of 4 rows high
    for slabs in (1, 3, 5, [4, 8, 16, 40, 18, 8, 2]):
        st = R.StreamedFusedMSE(cfg, w2o.shape[0], cuda, slabs=slabs) if isinstance(slabs, int) else \
            R.StreamedFusedMSE(cfg, w2o.shape[0], cuda, heights=[h for h in slabs[:-2]] + [slabs[-2] + slabs[-1]])
        for rep in range(4):                     # later calls reuse buffers / events, scene alternates
            loss0, grad0, img0, _ = ref[rep & 1]
            pin_i.zero_()
            loss, grad = st(ot, (w2o, w2o_b)[rep & 1], mat, light, cam, pin_t, pin_i)
            torch.cuda.synchronize()
            assert torch.equal(pin_i, img0.cpu())
            np.testing.assert_allclose(float(loss), float(loss0), rtol=1e-6)
            s_ = float(grad0.abs().max())
            assert float((grad - grad0).abs().max()) <= 1e-4 * s_
    with pytest.raises(ValueError):
        st(ot, w2o, mat, light, cam, target.cpu(), pin_i)       # not pinned
    with pytest.raises(ValueError):
        R.StreamedFusedMSE(cfg, w2o.shape[0], cuda, heights=[4, 8])          # does not cover the image


def test_sparse_upstream_gradient_early_out(cuda):
    """Reverse pass with two non-zero upstream pixels (optimize_brightness.py:51): CTAs without
    upstream gradient leave early; result == oracle and == the dense kernel path on the same
    gradient with a tiny value added everywhere (which disables the early-out)."""
    ps = oc.PackedScene.from_spec(scenes.optimize_brightness(n=128, samples=4, seed=5), camera_grad=0)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, cuda)
    dl = np.zeros((1, 128, 128, 3), dtype=np.float32)
    dl[0, 90, 85] = -1.0
    dl[0, 50, 90] = -1.0
    g_o = oc.render_backward(ps, dl)[0]
    g = R.render_backward(cfg, ot, w2o, mat, light, cam, torch.from_numpy(dl).to(cuda), None, jit).double().cpu().numpy()
    assert np.max(np.abs(g - g_o)) <= 1e-3 * np.max(np.abs(g_o))
    assert np.max(np.abs(g_o)) > 0
    g_zero = R.render_backward(cfg, ot, w2o, mat, light, cam, torch.zeros((1, 128, 128, 3), device=cuda), None, jit)
    assert float(g_zero.abs().max()) == 0.0


@pytest.mark.parametrize('name', ['C1_optimize_brightness', 'C3_match_mirror_square', 'C4_orbit_view0', 'S8', 'S1'])
def test_small_and_general_kernels_agree(name, cuda):
    """The small-scene kernel and the general kernel share the canonical routines and the
    sample summation order: masks, tmin AND pixels are bit-identical; gradients agree up to
    the reduction order."""
    from dataclasses import replace
    ps = oc.PackedScene.from_spec(CASES[name](), camera_grad=1)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, cuda)
    a = R.render_forward(replace(cfg, no_small=0), ot, w2o, mat, light, cam, jit, want_hit=True, want_tmin=True)
    b = R.render_forward(replace(cfg, no_small=1), ot, w2o, mat, light, cam, jit, want_hit=True, want_tmin=True)
    assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    assert torch.equal(a[0], b[0]), float((a[0] - b[0]).abs().max())
    target = torch.zeros_like(a[0])
    l0, g0, _, _ = R.render_fused_mse(replace(cfg, no_small=0), ot, w2o, mat, light, cam, target, None, jit)
    l1, g1, _, _ = R.render_fused_mse(replace(cfg, no_small=1), ot, w2o, mat, light, cam, target, None, jit)
    np.testing.assert_allclose(float(l0.sum()), float(l1.sum()), rtol=1e-6)
    assert float((g1 - g0).abs().max()) <= 1e-4 * float(g0.abs().max())


@pytest.mark.parametrize('samples,num_objects', [(16, 9), (16, 100), (12, 40), (32, 600)])
def test_fused_many_samples(samples, num_objects, cuda):
    """Fused mode beyond the 8 samples one thread of the general kernel holds: small scenes take the
    small-scene kernel (power-of-two S), everything else the general kernel's two-pass chunk loop
    (pass 0 sweeps + shades, pass 1 sweeps again + reverse pass) -- S = 16 with 100 shapes used to
    be refused with RRT_ERR_UNSUPPORTED."""
    ps = oc.PackedScene.from_spec(scenes.stress(n=20, num_objects=num_objects, samples=samples), camera_grad=1)
    img_o, _, _ = oc.render_forward(ps, want_aux=False)
    target = np.clip(img_o + 0.1, 0, 1).astype(np.float32)
    image_o, hit_o, loss_o, grad_o = oc.render_fused_mse(ps, target)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, cuda)
    loss, grad, image, hit = R.render_fused_mse(cfg, ot, w2o, mat, light, cam, torch.from_numpy(target).to(cuda),
                                                None, jit, want_image=True, want_hit=True)
    assert np.array_equal(hit.cpu().numpy().reshape(hit_o.shape), hit_o)
    np.testing.assert_allclose(image.cpu().numpy().reshape(image_o.shape), image_o, rtol=PIX_RTOL, atol=PIX_ATOL)
    np.testing.assert_allclose(float(loss), loss_o[0], rtol=1e-4)
    compare_grads(grad.cpu().numpy().astype(np.float64), grad_o[0], ps.N)


SHADOWED = 0x40000000


@pytest.mark.parametrize('case', ['phong', 'depth', 'general', 'many'])
def test_shadows_parity(case, cuda):
    """RRT_FLAG_SHADOWS (parity weakly pinned by the reference): winners AND
    the shadow flag bit-exact against the canonical C oracle, pixels / loss / gradients at the
    usual tolerances; forward, fused and backward (stored and re-swept winners)."""
    if case == 'many':                                   # > 512 objects: re-staged table chunks
        spec = scenes.stress(n=24, num_objects=700)
        spec['shadows'] = 1
    else:
        spec = scenes.shadow_scene(shader='depth' if case == 'depth' else 'phong', general=(case == 'general'))
    ps = oc.PackedScene.from_spec(spec, camera_grad=1)
    img_o, hit_o = check_forward_flags(ps, cuda)
    assert int(((hit_o >= 0) & ((hit_o & SHADOWED) != 0)).sum()) > 20
    rng = np.random.RandomState(3)
    target = np.clip(img_o + rng.normal(0, 0.1, img_o.shape), 0, 1).astype(np.float32)
    image_o, hit_f, loss_o, grad_o = oc.render_fused_mse(ps, target)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, cuda)
    assert cfg.shadows == 1
    loss, grad, image, hit = R.render_fused_mse(cfg, ot, w2o, mat, light, cam, torch.from_numpy(target).to(cuda),
                                                None, jit, want_image=True, want_hit=True)
    assert np.array_equal(hit.cpu().numpy().reshape(hit_f.shape), hit_f)
    np.testing.assert_allclose(image.cpu().numpy().reshape(image_o.shape), image_o, rtol=PIX_RTOL, atol=PIX_ATOL)
    np.testing.assert_allclose(float(loss), loss_o[0], rtol=1e-4)
    compare_grads(grad.cpu().numpy().astype(np.float64), grad_o[0], ps.N)
    # general kernel: packed (FFMA2) filter + scalar decision == scalar pass only (RRT_FLAG_SCALAR_SHADOWS)
    from dataclasses import replace
    for mode in (1, 2):
        h2 = R.render_forward(replace(cfg, no_small=1, shadows=mode), ot, w2o, mat, light, cam, jit, want_hit=True)[1]
        assert np.array_equal(h2.cpu().numpy().reshape(hit_o.shape), hit_o), mode
    dl = rng.normal(0, 1, img_o.shape).astype(np.float32)
    gb_o = oc.render_backward(ps, dl, hit_o)
    for stored in (True, False):
        h = torch.from_numpy(hit_o).to(cuda) if stored else None
        gb = R.render_backward(cfg, ot, w2o, mat, light, cam, torch.from_numpy(dl).to(cuda), h, jit)
        compare_grads(gb.cpu().numpy().astype(np.float64), gb_o[0], ps.N)


def check_forward_flags(ps, dev):
    img_o, hit_o, tmin_o = oc.render_forward(ps)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, dev)
    img, hit, tmin = R.render_forward(cfg, ot, w2o, mat, light, cam, jit, want_hit=True, want_tmin=True)
    assert np.array_equal(hit.cpu().numpy().reshape(hit_o.shape), hit_o), 'winners + shadow flags must be bit-exact'
    assert np.array_equal(tmin.cpu().numpy().reshape(tmin_o.shape).view(np.uint32), tmin_o.view(np.uint32))
    np.testing.assert_allclose(img.cpu().numpy().reshape(img_o.shape), img_o, rtol=PIX_RTOL, atol=PIX_ATOL)
    return img_o, hit_o


def _pathological_spec(n=48):
    """Objects at the edges of float32: overflow to inf / NaN discriminants, denormal
    discriminants, NaN and inf matrix entries, a zero matrix, a sphere behind the camera, an
    edge-on square.  Whatever IEEE says happens must happen
    identically in the kernels and in the canonical oracle (same operations, same order)."""
    from oracle import oracle_numpy as on
    spec = scenes.stress(n=n, num_objects=14, samples=4, seed=77)
    w = spec['w2o']                                       # [N,4,4] float32
    def ts(c, s_):
        return on.inverse(on.compose(on.translate(c), on.scale(s_)))[0]
    w[0] = ts((0.1, 0.1, 5.0), (1e-20, 1e-20, 1e-20))     # d' ~ 1e20: vn overflows, det = inf - inf
    w[1] = ts((1e18, 2e18, 3e19), (1e18, 1e18, 1e18))     # d' ~ 1e-18: denormal products, t ~ 3e19
    w[2] = ts((2.0, -1.5, -5.0), (1.2, 1.2, 1.2))         # behind the camera: negative t still wins
    w[3] = ts((0.5, 0.5, 3e-20), (1e-19, 1e-19, 1e-19))   # tiny sphere almost at the camera: o' ~ 1
    w[4] = w[4].copy(); w[4][0, 0] = np.nan
    w[5] = w[5].copy(); w[5][1, 3] = np.inf
    w[6] = np.zeros((4, 4), dtype=np.float32); w[6][3, 3] = 1
    w[7] = on.inverse(on.compose(on.translate((0.0, 0.0, 4.0)), on.rotate(90, (0., 1., 0.))))[0]   # edge-on square
    spec['obj_type'] = spec['obj_type'].copy(); spec['obj_type'][7] = on.SQUARE
    w[8] = ts((0.2, 3.0, 20.0), (3e38, 0.4, 0.4))         # 1/s denormal on one axis
    w[9] = ts((-0.4, 0.3, 6.0), (1e-38, 1e-38, 1e-38))    # 1/s = inf
    return spec


def test_pathological_objects_masks_bit_exact(cuda):
    with np.errstate(all='ignore'):
        ps = oc.PackedScene.from_spec(_pathological_spec(), camera_grad=0)
        img_o, hit_o, tmin_o = oc.render_forward(ps)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, cuda)
    img, hit, tmin = R.render_forward(cfg, ot, w2o, mat, light, cam, jit, want_hit=True, want_tmin=True)
    assert np.array_equal(hit.cpu().numpy().reshape(hit_o.shape), hit_o)
    assert np.array_equal(tmin.cpu().numpy().reshape(tmin_o.shape).view(np.uint32), tmin_o.view(np.uint32))
    assert len(np.unique(hit_o)) >= 6                      # several of the odd objects do win rays
    a, b = img.cpu().numpy().reshape(img_o.shape), img_o
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    np.testing.assert_allclose(a[fin], b[fin], rtol=PIX_RTOL, atol=PIX_ATOL)
    # culling must not change a bit either (general kernel; NaN-safe comparisons keep odd objects)
    from dataclasses import replace
    h2 = R.render_forward(replace(cfg, cull=1), ot, w2o, mat, light, cam, jit, want_hit=True)[1]
    assert torch.equal(h2, hit)


def test_record_table_is_bit_identical(cuda):
    """rrt_build_records + TMA staging vs records built per CTA: every output bit-identical
    (two table chunks, squares + general spheres, culling on and off)."""
    from dataclasses import replace
    ps = oc.PackedScene.from_spec(_mixed_many(), camera_grad=1)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, cuda)
    for cull in (0, 1):
        a = R.render_forward(replace(cfg, use_records=1, cull=cull), ot, w2o, mat, light, cam, jit, want_hit=True, want_tmin=True)
        b = R.render_forward(replace(cfg, use_records=0, cull=cull), ot, w2o, mat, light, cam, jit, want_hit=True, want_tmin=True)
        for x, y in zip(a, b):
            assert torch.equal(x, y)


@pytest.mark.parametrize('name', ['C1_optimize_brightness', 'C2_test_balls_depth', 'C3_match_mirror_square', 'C4_orbit_view1',
                                  'C5_stress_diag', 'many_general_chunked'])
def test_no_material_grad_flag_and_in_kernel_finalize(name, cuda):
    """RRT_FLAG_NO_MATERIAL_GRAD (geom_grad_only): d/d w2o and d/d camera.o2w as without the flag,
    material / light / look_at entries exactly zero.  rrt_scene.ticket (last CTA of a scene finalises
    inside the render kernel, one launch) against the separate finalize launch.  Fused and
    backward entry points, whichever kernel the fixture selects."""
    ps = oc.PackedScene.from_spec(CASES[name](), camera_grad=1)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, cuda)
    img, hit, _ = R.render_forward(cfg, ot, w2o, mat, light, cam, jit)
    target = (img * 0.5 + 0.1).contiguous()
    dl = torch.randn_like(img)
    N = ps.N

    def both(c):
        out = [R.render_backward(c, ot, w2o, mat, light, cam, dl, h, jit) for h in (hit, None)]
        if c.samples <= 8 or not c.no_small:
            out.append(R.render_fused_mse(c, ot, w2o, mat, light, cam, target, None, jit)[1])
        return out
    full = both(replace(cfg, use_ticket=0))
    for variant in (dict(use_ticket=1), dict(use_ticket=1, geom_grad_only=1), dict(use_ticket=0, geom_grad_only=1)):
        for g, g0 in zip(both(replace(cfg, **variant)), full):
            gw, gm, gl, gc = R.split_grad(g, N)
            gw0, gm0, gl0, gc0 = R.split_grad(g0, N)
            tol = 1e-5
            assert float((gw - gw0).abs().max()) <= tol * max(float(gw0.abs().max()), 1e-30)
            assert float((gc[:12] - gc0[:12]).abs().max()) <= tol * max(float(gc0[:12].abs().max()), 1e-30)
            if variant.get('geom_grad_only'):
                assert float(gm.abs().max()) == 0.0 and float(gl.abs().max()) == 0.0 and float(gc[12:].abs().max()) == 0.0
            else:
                assert float((gm - gm0).abs().max()) <= tol * max(float(gm0.abs().max()), 1e-30)
                assert float((gl - gl0).abs().max()) <= tol * max(float(gl0.abs().max()), 1e-30)
    # the scratch is left zero for the next call
    tk = R._ticket(cuda, 1)
    assert tk is not None and int(tk.abs().sum()) == 0


@pytest.mark.parametrize('mapping', [1, 2])
def test_many_small_scenes_persistent_ctas(mapping, cuda):
    """(Both thread mappings of the batch: 1 = one pixel per thread, 2 = one ray per thread; the
    single-scene launches it is compared with take the library's own choice.)
    The small-scene kernel's persistent CTAs walk contiguous ranges of work items that cross
    scene boundaries (here 3000 scenes of 2 items: every CTA serves several scenes, flushing its
    sums, reloading the tables and taking a per-scene ticket at each boundary).  Per-scene loss,
    gradient, image and masks must equal single-scene launches of the same scenes."""
    B, n = 3000, 8
    tb = W.orbit_tables(B // 2)
    rng = np.random.RandomState(5)
    w2o = tb['w2o'].copy()
    w2o[:, 0, 3] += rng.uniform(-0.3, 0.3, B).astype(np.float32)          # every scene differs
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    cfg = R.RenderConfig(n=n, samples=4, shader=tb['shader'], transpose=0, seed=11, camera_grad=1)
    ot, mat, light, cam, w2o_d = t(tb['obj_type']), t(tb['material']), t(tb['light']), t(tb['camera']), t(w2o)
    target = torch.rand((B, n, n, 3), device=cuda)
    loss, grad, image, hit = R.render_fused_mse(replace(cfg, pixel_threads=mapping), ot, w2o_d, mat, light, cam, target,
                                                want_image=True, want_hit=True)
    dl = torch.randn_like(image)
    gb = R.render_backward(cfg, ot, w2o_d, mat, light, cam, dl, None)
    for b in (0, 1, 2, 777, 1500, 2998, 2999):
        c1 = replace(cfg, scene_begin=b)                                   # keys the in-kernel jitter like the batch
        l1, g1, im1, h1 = R.render_fused_mse(c1, ot, w2o_d[b], mat, light, cam[b], target[b], want_image=True, want_hit=True)
        assert torch.equal(h1, hit[b]) and torch.equal(im1, image[b])
        np.testing.assert_allclose(float(l1), float(loss[b]), rtol=1e-6)
        assert float((g1 - grad[b]).abs().max()) <= 1e-5 * float(g1.abs().max())
        gb1 = R.render_backward(c1, ot, w2o_d[b], mat, light, cam[b], dl[b], None)
        assert float((gb1 - gb[b]).abs().max()) <= 1e-5 * float(gb1.abs().max())


def _prefilter_equal(ps, cuda, with_fused=True):
    """pre-filter sweep (default with a record table) vs RRT_FLAG_CANONICAL_SWEEP: every output bit."""
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, cuda)
    cfg = replace(cfg, no_small=1, use_records=1)
    a = R.render_forward(replace(cfg, canonical_sweep=0), ot, w2o, mat, light, cam, jit, want_hit=True, want_tmin=True)
    b = R.render_forward(replace(cfg, canonical_sweep=1), ot, w2o, mat, light, cam, jit, want_hit=True, want_tmin=True)
    assert torch.equal(a[1], b[1]), 'hit masks differ'
    assert torch.equal(a[2].view(torch.int32), b[2].view(torch.int32)), 'tmin differs'
    assert torch.equal(a[0].view(torch.int32), b[0].view(torch.int32)), 'image differs'
    return a


@pytest.mark.parametrize('seed', range(12))
def test_prefilter_sweep_is_bit_identical_random(seed, cuda):
    """Random scenes over a wide range of scales, distances, rotations and anisotropy (>= 64 objects so
    that the record table + pre-filter are active), spheres behind the camera, camera inside a sphere,
    a rotated / translated camera (orbit variant)."""
    from oracle import oracle_numpy as on
    rng = np.random.RandomState(1000 + seed)
    N = int(rng.choice([64, 100, 257, 600]))
    shapes = []
    for k in range(N):
        z = rng.uniform(-4, 30)
        c = (rng.uniform(-0.6, 0.6) * abs(z) - 0.1, rng.uniform(-0.6, 0.6) * abs(z) + 0.1, z)
        t = on.translate(c)
        if rng.rand() < 0.5:
            ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
            t = on.compose(t, on.rotate(rng.uniform(0, 360), ax))
        sc = 10.0 ** rng.uniform(-2.5, 1.2, 3) if rng.rand() < 0.5 else np.full(3, 10.0 ** rng.uniform(-2.5, 1.2))
        t = on.compose(t, on.scale(sc))
        shapes.append((on.SPHERE, t, scenes._mat(rng.uniform(0.1, 1, 3), 0.3, 0.7, 0.4, 50.)))
    cam = None
    if seed % 3 == 1:
        cam = on.compose(on.translate(rng.uniform(-1, 1, 3)), on.rotate(rng.uniform(-40, 40), (0, 1, 0)))
    spec = scenes.spec_from(40, 4, shapes, ((-1., -1., 2.), (0.961, 1., 0.87)), 'phong', cam=cam, seed=seed)
    ps = oc.PackedScene.from_spec(spec, camera_grad=0)
    img, hit, tmin = _prefilter_equal(ps, cuda)
    img_o, hit_o, tmin_o = oc.render_forward(ps)
    assert np.array_equal(hit.cpu().numpy().reshape(hit_o.shape), hit_o)
    assert np.array_equal(tmin.cpu().numpy().reshape(tmin_o.shape).view(np.int32), tmin_o.view(np.int32))


def test_prefilter_sweep_out_of_range_inputs(cuda):
    """Objects and rays outside the range the pre-filter's bound is proven for must fall back to the
    canonical arithmetic: huge / tiny / zero / NaN / inf matrices (always-pass rows), a camera whose
    rays have d_z <= 0 or |d_x/d_z| > 2^10 (CTA-wide canonical sweep), squares mixed in."""
    from oracle import oracle_numpy as on
    base = scenes.stress(n=48, num_objects=80)
    w = base['w2o'].copy()                                   # [N, 4, 4]
    w[3, :3] *= 3.0e7; w[4, :3] *= 1.0e-7; w[5, :3, :3] = 0.0; w[6, 0, 0] = np.nan; w[7, 1, 3] = np.inf; w[8, :3, 3] *= 1.0e8
    w[9, :3] *= 2.0e5; w[10, :3] *= 5.0e-6
    base['w2o'] = w
    base['obj_type'] = base['obj_type'].copy()
    base['obj_type'][70:] = on.SQUARE                        # chunk with squares: canonical mixed sweep
    for cam in (None, on.rotate(89.99, (0, 1, 0)), on.rotate(180, (0, 1, 0)), on.compose(on.translate((0, 0, 12)), on.rotate(90, (1, 0, 0)))):
        spec = dict(base)
        spec['cam_o2w'] = None if cam is None else np.asarray(cam[0], dtype=np.float32)
        ps = oc.PackedScene.from_spec(spec, camera_grad=0)
        _, hit, tmin = _prefilter_equal(ps, cuda)
        _, hit_o, tmin_o = oc.render_forward(ps)
        assert np.array_equal(hit.cpu().numpy().reshape(hit_o.shape), hit_o)
        assert np.array_equal(tmin.cpu().numpy().reshape(tmin_o.shape).view(np.int32), tmin_o.view(np.int32))


@pytest.mark.parametrize('general', [False, True])
def test_prefilter_sweep_full_size_slab(general, cuda):
    """BASELINE config 5 at full resolution (4096 x 4096, 1024 spheres; C5 and C5g), a 96-row slab in the
    middle of the image: pre-filter and canonical sweeps agree on every bit of hit_index, tmin, image."""
    tb = W.stress_tables(1024, general=general)
    t = lambda a: torch.from_numpy(a).to(cuda)
    args = (t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']), t(tb['camera']))
    cfg = R.RenderConfig(n=4096, samples=4, shader=tb['shader'], transpose=1, seed=4321, row_begin=2000, row_count=96)
    a = R.render_forward(cfg, *args, None, want_hit=True, want_tmin=True)
    b = R.render_forward(replace(cfg, canonical_sweep=1), *args, None, want_hit=True, want_tmin=True)
    assert torch.equal(a[1], b[1]) and torch.equal(a[2].view(torch.int32), b[2].view(torch.int32))
    assert torch.equal(a[0].view(torch.int32), b[0].view(torch.int32))
    assert int((a[1] >= 0).sum()) > 100000


@pytest.mark.parametrize('name', ['C3_match_mirror_square', 'C4_orbit_view0', 'C5_stress_diag', 'C5g_stress_general',
                                  'many_mixed_chunked'])
def test_deterministic_mode_is_bitwise_reproducible(name, cuda):
    """RRT_FLAG_DETERMINISTIC: loss and gradients are bit-identical from run to run (the reference's
    T.grad is, optimize.py:25; float atomics are not) and agree with the default mode to float32
    rounding.  Fused and backward entry points, with and without the in-kernel finalisation."""
    ps = oc.PackedScene.from_spec(CASES[name](), camera_grad=1)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, cuda)
    img, hit, _ = R.render_forward(cfg, ot, w2o, mat, light, cam, jit)
    target = (img * 0.5 + 0.1).contiguous()
    dl = torch.randn_like(img)
    fused_ok = cfg.samples <= 8 or not cfg.no_small
    for ticket in (1, 0):
        c = replace(cfg, deterministic=1, use_ticket=ticket)
        runs = []
        for rep in range(4):
            g_b = R.render_backward(c, ot, w2o, mat, light, cam, dl, None, jit)
            l_f, g_f = R.render_fused_mse(c, ot, w2o, mat, light, cam, target, None, jit)[:2] if fused_ok else (g_b[:1], g_b)
            runs.append((g_b.clone(), l_f.clone(), g_f.clone()))
        for r in runs[1:]:
            assert torch.equal(r[0], runs[0][0]) and torch.equal(r[1], runs[0][1]) and torch.equal(r[2], runs[0][2])
        g0 = R.render_backward(replace(cfg, use_ticket=ticket), ot, w2o, mat, light, cam, dl, None, jit)
        assert float((runs[0][0] - g0).abs().max()) <= 2e-5 * float(g0.abs().max())
        if fused_ok:
            l0, gf0 = R.render_fused_mse(replace(cfg, use_ticket=ticket), ot, w2o, mat, light, cam, target, None, jit)[:2]
            np.testing.assert_allclose(float(runs[0][1]), float(l0), rtol=1e-6)
            assert float((runs[0][2] - gf0).abs().max()) <= 2e-5 * float(gf0.abs().max())


def test_deterministic_mode_batch_and_streamed(cuda):
    """Deterministic mode on a scene batch (persistent small-scene CTAs crossing scene boundaries) and
    through StreamedFusedMSE (two kernel streams, one workspace each)."""
    B, n = 600, 16
    tb = W.orbit_tables(B // 2)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    cfg = R.RenderConfig(n=n, samples=4, shader=tb['shader'], transpose=0, seed=3, camera_grad=1, deterministic=1)
    args = (t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']), t(tb['camera']))
    target = torch.rand((B, n, n, 3), device=cuda)
    a = R.render_fused_mse(cfg, *args, target)
    for _ in range(3):
        b = R.render_fused_mse(cfg, *args, target)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    ref = R.render_fused_mse(replace(cfg, deterministic=0), *args, target)
    np.testing.assert_allclose(a[0].cpu().numpy(), ref[0].cpu().numpy(), rtol=1e-6)
    assert float((a[1] - ref[1]).abs().max()) <= 2e-5 * float(ref[1].abs().max())
    # streamed, 100 objects (record table + pre-filter), deterministic
    ps = oc.PackedScene.from_spec(scenes.stress(n=96, num_objects=100, samples=4, seed=11), camera_grad=0)
    c0, ot, w2o, mat, light, cam, _ = to_device(ps, cuda, with_jitter=False)
    c = R.RenderConfig(n=c0.n, samples=4, shader=c0.shader, transpose=c0.transpose, seed=5, deterministic=1)
    tgt = torch.rand((c.n, c.n, 3), device=cuda)
    pin_t, pin_i = tgt.cpu().pin_memory(), torch.empty((c.n, c.n, 3)).pin_memory()
    st = R.StreamedFusedMSE(c, w2o.shape[0], cuda, slabs=4)
    outs = []
    for _ in range(3):
        l, g = st(ot, w2o, mat, light, cam, pin_t, pin_i)
        torch.cuda.synchronize()
        outs.append((l.clone(), g.clone()))
    assert all(torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1]) for o in outs[1:])


@pytest.mark.parametrize('general', [False, True])
def test_mirror_bounce_parity(general, cuda):
    """RRT_FLAG_MIRROR (one reflection bounce; an extension, PARITY UNPINNED by the reference, pinned by
    the dense NumPy restatement, float64 autograd and finite differences in tests/test_oracle_golden.py):
    the kernels against the C oracle -- primary masks bit-exact, pixels rtol 1e-4 (the secondary hit
    of every reflected ray is decided in canonical float32 order on both sides), gradients <= 1e-3
    per block.  Forward, backward (stored and re-swept winners) and fused entry points, both kernels."""
    from helpers import reflectivity_of
    spec = scenes.mirror_scene(n=64, general=general)
    ps = oc.PackedScene.from_spec(spec, camera_grad=0)
    img_o, hit_o, hit2_o = oc.render_forward_secondary(ps)
    assert (hit2_o >= 0).mean() > 0.03                          # reflections are actually visible
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, cuda)
    refl = reflectivity_of(ps, cuda)
    img, hit, _ = R.render_forward(cfg, ot, w2o, mat, light, cam, jit, reflectivity=refl)
    assert np.array_equal(hit.cpu().numpy().reshape(hit_o.shape), hit_o)
    np.testing.assert_allclose(img.cpu().numpy().reshape(img_o.shape), img_o, rtol=PIX_RTOL, atol=PIX_ATOL)
    plain, _, _ = R.render_forward(cfg, ot, w2o, mat, light, cam, jit)
    assert float((plain - img).abs().max()) > 0.1               # the flag changes the picture
    rng = np.random.RandomState(3)
    dl = rng.normal(size=img_o[0].shape).astype(np.float32)
    grad_o = oc.render_backward(ps, dl, hit_o)
    for stored in (hit, None):
        g = R.render_backward(cfg, ot, w2o, mat, light, cam, torch.from_numpy(dl).to(cuda), stored, jit, reflectivity=refl)
        compare_grads(g.cpu().numpy().astype(np.float64), grad_o[0], ps.N)
    target = np.clip(img_o[0] + rng.normal(0, 0.1, img_o[0].shape), 0, 1).astype(np.float32)
    image_o, _, loss_o, gradf_o = oc.render_fused_mse(ps, target)
    loss, grad, image, _ = R.render_fused_mse(cfg, ot, w2o, mat, light, cam, torch.from_numpy(target).to(cuda), None, jit,
                                              want_image=True, reflectivity=refl)
    np.testing.assert_allclose(image.cpu().numpy().reshape(image_o.shape), image_o, rtol=PIX_RTOL, atol=PIX_ATOL)
    np.testing.assert_allclose(float(loss), loss_o[0], rtol=1e-4)
    compare_grads(grad.cpu().numpy().astype(np.float64), gradf_o[0], ps.N)
    # deterministic mode covers the bounce, too
    c = replace(cfg, deterministic=1)
    a = R.render_fused_mse(c, ot, w2o, mat, light, cam, torch.from_numpy(target).to(cuda), None, jit, reflectivity=refl)
    b = R.render_fused_mse(c, ot, w2o, mat, light, cam, torch.from_numpy(target).to(cuda), None, jit, reflectivity=refl)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_non_integer_shininess_nan_propagates_like_the_reference(cuda):
    """shader.py:45 raises (rm . look_at) to `shininess` with Theano's pow: a negative base with a
    non-integer exponent is NaN, and T.clip passes NaN on -- so does the kernel (its clip is
    switch-based, not fmin/fmax) and the oracle; pixels agree including where they are NaN."""
    spec = scenes.optimize_brightness(n=64)
    spec['material'] = spec['material'].copy()
    spec['material'][:, 3] = 50.5                              # non-integer shininess
    ps = oc.PackedScene.from_spec(spec, camera_grad=0)
    img_o, hit_o, _ = oc.render_forward(ps)
    assert np.isnan(img_o).any() and np.isfinite(img_o).any()
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, cuda)
    img, hit, _ = R.render_forward(cfg, ot, w2o, mat, light, cam, jit)
    assert np.array_equal(hit.cpu().numpy().reshape(hit_o.shape), hit_o)
    got = img.cpu().numpy().reshape(img_o.shape)
    assert np.array_equal(np.isnan(got), np.isnan(img_o))
    np.testing.assert_allclose(got, img_o, rtol=1e-3, atol=1e-5, equal_nan=True)


@pytest.mark.parametrize('name', ['C1_optimize_brightness', 'C2_test_balls_depth', 'C3_match_mirror_square', 'C4_orbit_view1',
                                  'S1', 'S2', 'ragged_n1', 'ragged_n5'])
@pytest.mark.parametrize('geom', [0, 1])
def test_pixel_and_ray_thread_mappings_agree(name, geom, cuda):
    """The two thread mappings of the small-scene kernel (one pixel per thread with its S samples in
    registers, 8 x 4 pixel tiles per warp / one ray per thread, samples combined by shuffles) run the
    same device routines in the same sample order: masks, tmin and pixels are bit-identical, loss and
    gradients agree up to the reduction order.  Also as row slabs whose height is not a multiple of
    the tile height, and in deterministic mode (where the two mappings must agree to the last bits
    the fixed-point sums keep)."""
    ps = oc.PackedScene.from_spec(CASES[name](), camera_grad=1)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, cuda)
    cfg = replace(cfg, no_small=0, geom_grad_only=geom)
    pix, ray = replace(cfg, pixel_threads=1), replace(cfg, pixel_threads=2)
    a = R.render_forward(pix, ot, w2o, mat, light, cam, jit, want_hit=True, want_tmin=True)
    b = R.render_forward(ray, ot, w2o, mat, light, cam, jit, want_hit=True, want_tmin=True)
    assert torch.equal(a[1], b[1]) and torch.equal(a[2].view(torch.int32), b[2].view(torch.int32))
    assert torch.equal(a[0].view(torch.int32), b[0].view(torch.int32)), float((a[0] - b[0]).abs().max())
    target = (a[0] * 0.5 + 0.1).contiguous()
    for det in (0, 1):
        la, ga, ia, ha = R.render_fused_mse(replace(pix, deterministic=det), ot, w2o, mat, light, cam, target, None, jit,
                                            want_image=True, want_hit=True)
        lb, gb, ib, hb = R.render_fused_mse(replace(ray, deterministic=det), ot, w2o, mat, light, cam, target, None, jit,
                                            want_image=True, want_hit=True)
        assert torch.equal(ha, hb) and torch.equal(ia.view(torch.int32), ib.view(torch.int32))
        np.testing.assert_allclose(float(la.sum()), float(lb.sum()), rtol=1e-6)
        assert float(gb.abs().max()) > 0 or ps.n == 1
        assert float((ga - gb).abs().max()) <= (2e-5 if det else 1e-4) * max(float(gb.abs().max()), 1e-30)
    # the reverse-only entry point, with stored and with re-swept winners, dense and sparse upstream gradients
    rng = np.random.RandomState(4)
    dl = torch.from_numpy(rng.normal(size=tuple(a[0].shape)).astype(np.float32)).to(cuda)
    sparse = torch.zeros_like(dl)
    sparse.view(-1, 3)[:: max(1, ps.n * ps.n // 3)] = 1.0
    for up in (dl, sparse):
        for stored in (a[1], None):
            ga = R.render_backward(pix, ot, w2o, mat, light, cam, up, stored, jit)
            gb = R.render_backward(ray, ot, w2o, mat, light, cam, up, stored, jit)
            assert float((ga - gb).abs().max()) <= 1e-4 * max(float(gb.abs().max()), 1e-30)
    # a row slab that starts inside a tile row and is not a multiple of 4 rows high, in-kernel jitter
    if ps.n >= 5:
        rb, rc = 1, min(ps.n - 1, 7)
        sp, sr = replace(pix, row_begin=rb, row_count=rc, seed=77), replace(ray, row_begin=rb, row_count=rc, seed=77)
        a = R.render_forward(sp, ot, w2o, mat, light, cam, None, want_hit=True, want_tmin=True)
        b = R.render_forward(sr, ot, w2o, mat, light, cam, None, want_hit=True, want_tmin=True)
        assert torch.equal(a[1], b[1]) and torch.equal(a[2].view(torch.int32), b[2].view(torch.int32))
        assert torch.equal(a[0].view(torch.int32), b[0].view(torch.int32))
        tgt = torch.zeros_like(a[0])
        la, ga, _, _ = R.render_fused_mse(sp, ot, w2o, mat, light, cam, tgt)
        lb, gb, _, _ = R.render_fused_mse(sr, ot, w2o, mat, light, cam, tgt)
        np.testing.assert_allclose(float(la.sum()), float(lb.sum()), rtol=1e-6)
        assert float((ga - gb).abs().max()) <= 1e-4 * max(float(gb.abs().max()), 1e-30)


@pytest.mark.parametrize('name', ['C1_optimize_brightness', 'C3_match_mirror_square', 'C4_orbit_view0', 'C5_stress_diag', 'S1', 'S2',
                                  'S8', 'ragged_n5'])
@pytest.mark.parametrize('dense', [False, True])
def test_linear_cost_equals_forward_plus_backward(name, dense, cuda):
    """RRT_FLAG_LINEAR_COST: the fused entry point with a weight image W (optimize_brightness.py:51 is W = -1 at
    two pixels) == the forward render (image bit for bit, loss = sum(W * image)) + the reverse-pass entry point with
    dL/dimage = W (whose parity with the oracle is established above) -- on every kernel / thread mapping."""
    ps = oc.PackedScene.from_spec(CASES[name](), camera_grad=1)
    cfg, ot, w2o, mat, light, cam, jit = to_device(ps, cuda)
    img, hit, _ = R.render_forward(cfg, ot, w2o, mat, light, cam, jit, want_hit=True)
    rng = np.random.RandomState(11)
    if dense:
        Wt = torch.from_numpy(rng.normal(size=tuple(img.shape)).astype(np.float32)).to(cuda)
    else:
        Wt = torch.zeros_like(img)
        flat = Wt.view(-1, 3)
        lit = torch.nonzero((hit.reshape(ps.samples, -1) >= 0).any(0)).reshape(-1)       # pixels that see an object
        pick_ = lit[torch.from_numpy(rng.choice(len(lit), size=min(3, len(lit)), replace=False)).to(cuda)] if len(lit) else lit
        flat[pick_] = -1.0
    cw = (1.0, 0.5, 2.0)
    loss, grad, image, hit2 = R.render_fused_mse(replace(cfg, linear_cost=1), ot, w2o, mat, light, cam, Wt, cw, jit,
                                                 want_image=True, want_hit=True)
    assert torch.equal(hit2, hit) and torch.equal(image.view(torch.int32), img.view(torch.int32))
    cwt = torch.tensor(cw, device=cuda)
    expect = float((Wt.double() * img.double() * cwt.double()).sum())
    assert abs(float(loss.sum()) - expect) <= 1e-5 * max(1.0, float((Wt.abs() * img).sum()))
    gb = R.render_backward(cfg, ot, w2o, mat, light, cam, (Wt * cwt).contiguous(), None, jit)
    scale_ = max(float(gb.abs().max()), 1e-30)
    assert float((grad - gb).abs().max()) <= 2e-4 * scale_
    if len(torch.nonzero(Wt)) and ps.n > 1:
        assert float(gb.abs().max()) > 0
