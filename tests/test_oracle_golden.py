"""CPU tests that PIN the oracle -- no GPU needed.

  (1) the reference's own unit tests, test/test_transform.py:9-41 (3 known answers)
  (2) orbit_experiments/orbit_dataset.npz[99] (+ orbit_target.npz): forward, orbit variant
  (3) output/0.jpg: forward, root variant (match_mirror.py first frame), orientation
  (4) C oracle (canonical order) vs NumPy oracle (reference-structured): masks, pixels
  (5) primary rays: C oracle bit-exact vs the NumPy restatement of Camera.make_rays
  (6) gradient oracle: closed form == float64 autograd; finite differences
"""
import os

import numpy as np
import pytest
import torch

from oracle import oracle_c as oc, oracle_grad as og, oracle_numpy as on, scenes

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


# ---- (1) test/test_transform.py known answers ------------------------------------
def test_rotate_kat():
    m, _ = on.rotate(20, (0, 0, 1))
    assert np.all(np.isclose(m, np.array([[0.93969262, -0.34202015, 0., 0.], [0.34202015, 0.93969262, 0., 0.],
                                          [0., 0., 1., 0.], [0., 0., 0., 1.]])))


def test_composition_kat():
    m, _ = on.compose(on.translate((4, 5, 6)), on.rotate(20, (0, 0, 1)))
    assert np.all(np.isclose(m, np.array([[0.93969262, -0.34202015, 0., 4.], [0.34202015, 0.93969262, 0., 5.],
                                          [0., 0., 1., 6.], [0., 0., 0., 1.]])))


def test_apply_kat():
    rays = np.tile(np.array([0, 1, 0], dtype=np.float32), (10, 10, 1))
    m, _ = on.compose(on.translate((4, 5, 6)), on.rotate(90, (0, 0, 1)))
    o, r = on.apply_rayfield(m, (1, 0, 0), rays)
    assert np.all(np.isclose(o, [4, 6, 6]))
    assert np.all(np.isclose(r, np.tile([-1, 0, 0], (10, 10, 1)), atol=1e-6))


def test_apply_is_a_spatial_transpose():
    """transform.py:46: rays'[a,b] = m[:3,:3] @ rays[b,a] (test_apply cannot see it)."""
    rng = np.random.RandomState(0)
    rays = rng.normal(size=(5, 5, 3)).astype(np.float32)
    m, _ = on.rotate(30, (0, 1, 0))
    _, r = on.apply_rayfield(m, (0, 0, 0), rays)
    np.testing.assert_allclose(r[1, 3], m[:3, :3] @ rays[3, 1], rtol=1e-5, atol=1e-6)


# ---- (2) orbit golden render -----------------------------------------------------
@pytest.mark.parametrize('view', [0, 1])
@pytest.mark.parametrize('which', ['numpy', 'c'])
def test_orbit_sample99(view, which):
    g = np.load(os.path.join(GOLD, 'orbit_sample99.npz'))
    spec = scenes.orbit(g['centre'], view, seed=7)     # the reference's jitter was unseeded
    if which == 'numpy':
        img, idx, _ = on.render(spec)
    else:
        im, hit, _ = oc.render_forward(oc.PackedScene.from_spec(spec))
        img, idx = im[0].astype(np.float64), hit[0]
    u8 = (img * 255).astype(np.uint8)                  # planet_orbit.py:61
    ref = g['views'][view]
    diff = np.abs(u8.astype(int) - ref.astype(int))
    assert diff.mean() < 0.2                            # survey probe: 0.10-0.14
    assert (diff.max(2) > 8).sum() < 60                 # only AA-jittered silhouette pixels
    assert ((ref.sum(2) > 0) != (idx >= 0).any(0)).sum() <= 20


# ---- (3) match_mirror first frame ------------------------------------------------
def test_match_mirror_frame0_and_orientation():
    from scipy.ndimage import maximum_filter, minimum_filter
    ref = np.load(os.path.join(GOLD, 'match_mirror_frame0.npy')).astype(int)
    img, idx, _ = on.render(scenes.match_mirror(seed=3))
    u8 = np.clip(img * 255, 0, 255).astype(int)
    assert np.abs(u8 - ref).mean() < 2.5                              # JPEG noise + AA jitter
    assert np.abs(u8.transpose(1, 0, 2) - ref).mean() > 30           # the transposed image does NOT match
    same = (idx == idx[0]).all(0)
    k0 = idx[0]
    interior = same & (maximum_filter(k0, 5) == minimum_filter(k0, 5)) & minimum_filter(same.astype(np.uint8), 5).astype(bool)
    assert interior.sum() > 10000
    d = np.abs(u8 - ref)[interior]
    assert d.max() <= 20 and d.mean() < 1.5                           # survey probe: 17 / 1.17


# ---- (4) canonical C oracle vs reference-structured NumPy oracle ------------------
CASES = {
    'C1': lambda: scenes.optimize_brightness(),
    'C2': lambda: scenes.test_balls(),
    'C3': lambda: scenes.match_mirror(),
    'C4': lambda: scenes.orbit((3.83, -8.14, 32), 0),
    'C5': lambda: scenes.stress(n=96, num_objects=48),
    'C5g': lambda: scenes.stress(n=96, num_objects=48, general=True),
}


@pytest.mark.parametrize('name', list(CASES))
def test_c_oracle_vs_numpy_oracle(name):
    spec = CASES[name]()
    img_n, idx_n, tm_n = on.render(spec)
    img_c, idx_c, tm_c = oc.render_forward(oc.PackedScene.from_spec(spec))
    mism = int((idx_n != idx_c[0]).sum())
    # FMA (canonical) vs plain NumPy rounding only moves silhouette rays
    assert mism <= max(4, idx_n.size // 10000), mism
    agree = (idx_n == idx_c[0]).all(0)
    # float32 det is ill-conditioned at grazing hits of far, small spheres (C5): bound loosely there
    tol = 5e-3 if name.startswith('C5') else 5e-5
    assert np.abs(img_n - img_c[0])[agree].max() < tol


# ---- (5) primary rays --------------------------------------------------------------
@pytest.mark.parametrize('n', [1, 2, 7, 64, 333])
def test_primary_rays_bit_exact(n):
    S = 4
    jx, jy = on.draw_jitter(n, S, np.random.RandomState(5))
    spec = dict(scenes.test_balls(n=n, samples=S))
    spec['jitter_x'], spec['jitter_y'] = jx, jy
    rays = oc.primary_rays(oc.PackedScene.from_spec(spec))
    ref = np.zeros_like(rays)
    for s in range(S):
        _, r = on.make_rays(n, n, (jx[:, :, s] + np.float32(s)) / np.float32(S), (jy[:, :, s] + np.float32(s)) / np.float32(S))
        ref[:, :, s, :] = r.transpose(1, 0, 2)          # root variant: pixel (a,b) <- ray [b,a]
    assert np.array_equal(rays.view(np.uint32), ref.view(np.uint32))


def test_rng_is_uniform_and_deterministic():
    v = np.array([oc.rng_value(4321, 0, p, s, a) for p in range(2000) for s in range(4) for a in range(2)])
    assert v.min() >= 0 and v.max() < 1
    assert abs(v.mean() - 0.5) < 0.02 and abs(v.var() - 1 / 12) < 0.01
    assert oc.rng_value(4321, 0, 17, 2, 1) == oc.rng_value(4321, 0, 17, 2, 1)
    assert oc.rng_value(4321, 0, 17, 2, 1) != oc.rng_value(4322, 0, 17, 2, 1)


# ---- (6) gradients -------------------------------------------------------------------
def _drop_inexact_winners(spec, idx):
    """Winners whose float32 det is > 0 but whose exact (float64) det is <= 0 (edge
    rays) would be sqrt(negative) in a float64 hit record: mark them background."""
    idx = idx.copy()
    R = og.effective_rays(spec).numpy()
    cam = np.eye(4) if spec.get('cam_o2w') is None else np.asarray(spec['cam_o2w'], dtype=np.float64)
    for k in range(len(spec['obj_type'])):
        if spec['obj_type'][k] != on.SPHERE:
            continue
        A = np.asarray(spec['w2o'][k], dtype=np.float64)
        o = A[:3, :3] @ cam[:3, 3] + A[:3, 3]
        d = R @ cam[:3, :3].T @ A[:3, :3].T
        pd, vn = d @ o, (d * d).sum(-1)
        det = pd * pd - vn * (o @ o - 1.0)
        idx[(idx == k) & ~(det > 1e-12)] = -1
    return idx


def _grad_blocks(spec, f64_record):
    ps = oc.PackedScene.from_spec(spec, camera_grad=1)
    img_c, idx_c, _ = oc.render_forward(ps)
    hit = _drop_inexact_winners(spec, idx_c[0]) if f64_record else idx_c[0]
    rng = np.random.RandomState(1)
    target = np.clip(img_c[0] + rng.normal(0, 0.1, img_c[0].shape), 0, 1).astype(np.float32)
    dl = (2 * (img_c[0] - target)).astype(np.float32)
    oc.lib().orc_set_f64_record(1 if f64_record else 0)
    try:
        grad = oc.render_backward(ps, dl, hit[None])
    finally:
        oc.lib().orc_set_f64_record(0)
    dlt = torch.from_numpy(dl).double()
    L, _, g = og.gradients(spec, hit, lambda im: (im * dlt).sum())
    gc = oc.split_grad(grad[0], ps.N)
    out = {}
    for k in ('w2o', 'material', 'light_dir', 'light_int', 'cam_o2w', 'look_at'):
        ref, a = g[k], gc[k]
        if k in ('w2o',):
            ref = ref[:, :3, :]
        if k == 'cam_o2w':
            ref = ref[:3, :]
        if k == 'material':                         # shininess slot: autograd oracle detaches it
            ref, a = np.delete(ref, 3, 1), np.delete(a, 3, 1)
        s = np.max(np.abs(ref))
        out[k] = float(np.max(np.abs(a - ref)) / s) if s > 0 else float(np.max(np.abs(a)))
    return out


@pytest.mark.parametrize('name', list(CASES))
def test_closed_form_equals_autograd(name):
    """Closed-form reverse pass (oracle_c, hit record in double) == float64 autograd."""
    errs = _grad_blocks(CASES[name](), f64_record=True)
    for k, e in errs.items():
        assert e < 2e-5, (k, e)


@pytest.mark.parametrize('name', ['C1', 'C2', 'C3', 'C4'])
def test_float32_record_conditioning(name):
    """Canonical float32 hit record vs exact float64 geometry: the residual is float32
    CONDITIONING of det at grazing hits (1/sqrt(det)), inherent to floatX=float32 -- it
    bounds how well ANY float32 renderer can match exact gradients, and is why kernel
    parity is measured against the float32-record oracle."""
    errs = _grad_blocks(CASES[name](), f64_record=False)
    for k, e in errs.items():
        assert e < 1e-2, (k, e)


def test_autograd_oracle_forward_matches_numpy_oracle():
    for name in ('C1', 'C3', 'C4'):
        spec = CASES[name]()
        img, idx, _ = on.render(spec)
        params = og.leaf_params(spec)
        with torch.no_grad():
            im2 = og.forward(spec, params, idx).numpy()
        # float32 geometry (numpy oracle) vs float64 geometry (autograd oracle)
        assert np.abs(img - im2).max() < 2e-3 and np.abs(img - im2).mean() < 1e-5


def test_autograd_oracle_finite_differences():
    """Central differences of the mask-constant forward (the semantics T.grad has)."""
    spec = scenes.optimize_brightness(n=48)
    _, idx, _ = on.render(spec)
    w = torch.from_numpy(np.random.RandomState(3).normal(size=(48, 48, 3)))
    loss_fn = lambda im: (im * w).sum()
    _, _, g = og.gradients(spec, idx, loss_fn)

    def f(name, index, eps):
        p = og.leaf_params(spec)
        with torch.no_grad():
            p[name][index] += eps
            return float(loss_fn(og.forward(spec, p, idx)))
    for name, index in (('w2o', (0, 0, 3)), ('w2o', (1, 1, 2)), ('w2o', (1, 2, 2)), ('material', (0, 1)),
                        ('material', (1, 5)), ('light_dir', (0,)), ('light_int', (2,)), ('look_at', (1,))):
        # geometry entries see 1/sqrt(det) at grazing rays (large higher derivatives): small step
        eps = 1e-8 if name == 'w2o' else 1e-6
        fd = (f(name, index, eps) - f(name, index, -eps)) / (2 * eps)
        assert abs(fd - g[name][index]) <= 1e-4 * max(1.0, abs(fd)), (name, index, fd, g[name][index])


def test_backward_entry_matches_fused():
    spec = CASES['C3']()
    ps = oc.PackedScene.from_spec(spec, camera_grad=1)
    img, hit, _ = oc.render_forward(ps)
    target = np.ascontiguousarray(img[0][:, ::-1, :])
    image, _, loss, grad = oc.render_fused_mse(ps, target)
    gb = oc.render_backward(ps, 2 * (image[0] - target), hit)
    np.testing.assert_allclose(gb, grad, rtol=1e-6, atol=1e-6)   # dl_dimage is rounded to float32
    # row slabs: gradients and losses add up
    g2, l2 = 0, 0
    for rb, rc in ((0, 50), (50, 78)):
        _, _, l, g = oc.render_fused_mse(ps.slab(rb, rc), target[rb:rb + rc])
        g2, l2 = g2 + g, l2 + l
    np.testing.assert_allclose(g2, grad, rtol=1e-9, atol=1e-7)
    np.testing.assert_allclose(l2, loss, rtol=1e-12)


# ---- (8) shadows: PARITY WEAKLY PINNED -- the reference's call site is
# commented out (scene.py:41-45) and its helper broken (shape.py:100-106); what is pinned is
# the formula of Sphere.shadow (shape.py:85-97): dense NumPy restatement vs canonical C.
SHADOWED = 0x40000000


@pytest.mark.parametrize('general', [False, True])
@pytest.mark.parametrize('shader', ['phong', 'depth'])
def test_shadows_c_oracle_vs_numpy_oracle(shader, general):
    spec = scenes.shadow_scene(shader=shader, general=general)
    img_n, idx_n, _ = on.render(spec)
    img_c, idx_c, _ = oc.render_forward(oc.PackedScene.from_spec(spec))
    nshadow = int(((idx_c[0] >= 0) & ((idx_c[0] & SHADOWED) != 0)).sum())
    assert nshadow > 200, nshadow                       # the scene really has shadows ...
    lit = int(((idx_c[0] >= 0) & ((idx_c[0] & SHADOWED) == 0)).sum())
    assert lit > 2000
    mism = int((idx_n != idx_c[0]).sum())
    assert mism <= 6, mism                              # float32-FMA vs float64 shadow edge rays
    agree = (idx_n == idx_c[0]).all(0)
    assert np.abs(img_n - img_c[0])[agree].max() < 5e-5
    # the flag off: nothing is shadowed and the winners are the same
    spec0 = dict(spec, shadows=0)
    _, idx0, _ = oc.render_forward(oc.PackedScene.from_spec(spec0))
    assert np.array_equal(idx0[0], np.where(idx_c[0] >= 0, idx_c[0] & ~SHADOWED, idx_c[0]))


def test_shadows_gradient_closed_form_equals_autograd():
    """Shadowed rays carry no gradient (masks are constants): closed form on the remaining
    winners == float64 autograd with the shadowed winners removed."""
    spec = scenes.shadow_scene(n=48)
    ps = oc.PackedScene.from_spec(spec, camera_grad=1)
    img_c, idx_c, _ = oc.render_forward(ps)
    hit_flagged = idx_c[0]
    lit_only = np.where((hit_flagged >= 0) & ((hit_flagged & SHADOWED) != 0), -1, hit_flagged)
    lit_only = _drop_inexact_winners(spec, lit_only)
    flagged = np.where(lit_only < 0, np.where(hit_flagged >= 0, hit_flagged | SHADOWED, -1), lit_only)
    rng = np.random.RandomState(2)
    dl = rng.normal(0, 1, img_c[0].shape).astype(np.float32)
    oc.lib().orc_set_f64_record(1)
    try:
        grad = oc.render_backward(ps, dl, flagged[None])       # stored winners incl. shadow flags
    finally:
        oc.lib().orc_set_f64_record(0)
    dlt = torch.from_numpy(dl).double()
    _, _, g = og.gradients(spec, lit_only, lambda im: (im * dlt).sum())
    gc = oc.split_grad(grad[0], ps.N)
    ref, a = g['w2o'][:, :3, :], gc['w2o']
    assert np.max(np.abs(a - ref)) <= 2e-5 * np.max(np.abs(ref))
    ref, a = g['light_dir'], gc['light_dir']
    assert np.max(np.abs(a - ref)) <= 2e-5 * np.max(np.abs(ref))
    # fused entry (sweeps + shadow test itself) == backward entry on its own flagged winners
    target = np.clip(img_c[0] + 0.1, 0, 1).astype(np.float32)
    image_f, hit_f, loss_f, grad_f = oc.render_fused_mse(ps, target)
    assert np.array_equal(hit_f[0], hit_flagged)
    grad_b = oc.render_backward(ps, (2 * (image_f[0] - target)).astype(np.float32), hit_f)
    assert np.max(np.abs(grad_f - grad_b)) <= 1e-9 * np.max(np.abs(grad_b))


# ---- DepthMapShader pinned to the reference's own depth renders (shader.py:14-20)
def _depth_spec(centres, seed=0):
    m1 = scenes._mat((0.2, 0.9, 0.4), 0.3, 0.7, 0.5, 50.)
    shapes = [(on.SPHERE, on.translate(c), m1) for c in centres]
    return scenes.spec_from(32, 4, shapes, ((-1., -1., 2.), (0.961, 1., 0.87)), 'depth', max_depth=6.1, seed=seed)


def _fit_depth_artefact(tgt_u8, init, iters=160, lr=2e-3):
    """Gradient descent on the sphere centres through the C oracle's fused forward + reverse pass.
    The artefacts were written by util.draw -> scipy.misc.imsave (generate_data.py:44-52), which
    rescales by the image maximum: u8 = round(255 * render / max(render)); the scale is re-derived
    from the current render on every step."""
    tgt = tgt_u8.astype(np.float64) / 255.0
    c = np.array(init, dtype=np.float64)
    for _ in range(iters):
        ps = oc.PackedScene.from_spec(_depth_spec(c))
        img, _, _ = oc.render_forward(ps, want_aux=False)
        k = 1.0 / float(img[0][:, :, 0].max())
        t3 = np.zeros((1, 32, 32, 3), np.float32)
        t3[0, :, :, 0] = tgt / k
        _, _, _, grad = oc.render_fused_mse(ps, t3, channel_weight=np.array([1, 0, 0], np.float32))
        c -= lr * (-oc.split_grad(grad[0], len(c))['w2o'][:, :, 3] * k * k)     # w2o = translate(-c): d/dc = -d/db
    return c


@pytest.mark.parametrize('which', ['15.jpg', 'example.png'])
def test_depth_shader_pinned_to_reference_depth_renders(which):
    """DepthMapShader(6.1) (shader.py:14-20, BASELINE config 2) against the two depth maps the reference
    ships: 15.jpg (test_balls.py:17: two unit spheres of generate_data.py's dataset, JPEG) and
    example.png (test_1ball.py:17: one sphere, lossless).  Their sphere centres are unknown
    (rand() in generate_data.py:39-40), so they are recovered by gradient descent through the oracle
    from a coarse grid estimate; the oracle must then reproduce the artefact.
      15.jpg: every pixel away from silhouettes to the JPEG floor (<= 8/255), mean |diff| 1.1/255,
        mismatches confined to silhouette pixels (the reference's AA jitter is unseeded), both centres
        inside the box the generator draws from ([-2,2]^2 x [4,6]).
      example.png: mean |diff| 1.7/255 with a unit sphere; its lower rim is up to 18/255 brighter than
        any unit sphere renders (provenance unknown: an older renderer or a non-unit scale, as in
        test_1ball.py's own model), so only the weaker bound is asserted there."""
    if which == '15.jpg':
        ref = np.load(os.path.join(GOLD, 'balls_15.npy'))
        init, interior_tol, max_out = [(-1.0, 0.0, 4.5), (1.0, -1.0, 5.0)], 8, 60
    else:
        ref = np.load(os.path.join(GOLD, 'example_png.npy'))
        init, interior_tol, max_out = [(1.5, 0.5, 4.5)], 20, 40
    c = _fit_depth_artefact(ref, init)
    assert np.all(np.abs(c[:, :2]) <= 2.0) and np.all((c[:, 2] >= 4.0) & (c[:, 2] <= 6.0)), c
    spec = _depth_spec(c)
    img_n, hit_n, _ = on.render(spec)
    ps = oc.PackedScene.from_spec(spec)
    im, h, _ = oc.render_forward(ps)
    for name, img in (('numpy', img_n[:, :, 0]), ('c', im[0][:, :, 0])):
        u8 = np.round(255.0 * img / img.max())
        d = np.abs(u8 - ref.astype(np.float64))
        assert d.mean() < 2.0, (which, name, d.mean())
        # pixels whose 4 samples agree in OUR render and whose 8 neighbours do too are interior or
        # background: they must match to the quantisation floor
        uniform = np.all(h[0] == h[0][0:1], axis=0)
        pad = np.pad(uniform, 1, mode='edge')
        inner = np.ones_like(uniform)
        for dy in (0, 1, 2):
            for dx in (0, 1, 2):
                inner &= pad[dy:dy + 32, dx:dx + 32]
        assert inner.sum() > 700
        assert d[inner].max() <= interior_tol, (which, name, d[inner].max())
        assert (d > 8).sum() <= max_out, (which, name, int((d > 8).sum()))


# ---- mirror bounce (RRT_FLAG_MIRROR): an EXTENSION with no reference at all -- match_mirror.py:40,45
# matches an image to its left-right flip, the hook would be scene.py:41-45 / shader.py:43-45 -- so
# PARITY IS UNPINNED by the reference.  Pinned by: dense NumPy restatement == canonical C (masks and
# images), closed-form reverse pass == float64 autograd, autograd == central finite differences.
@pytest.mark.parametrize('general', [False, True])
def test_mirror_c_oracle_vs_numpy_oracle(general):
    spec = scenes.mirror_scene(n=64, general=general)
    img_n, hit_n, _, hit2_n = on.render(spec, return_aux='mirror')
    ps = oc.PackedScene.from_spec(spec)
    img_c, hit_c, hit2_c = oc.render_forward_secondary(ps)
    assert (hit2_c >= 0).mean() > 0.03 and set(np.unique(hit2_c)) >= {-1, 0, 1, 2, 3}     # every object is seen in some mirror
    assert (hit_n != hit_c[0]).sum() <= 2 and (hit2_n != hit2_c[0]).sum() <= 6            # FMA vs no-FMA edge rays
    d = np.abs(img_n - img_c[0])
    assert (d > 1e-4).sum() <= 24 and np.median(d) < 1e-6
    # and the flag is not a no-op, nor does it touch rays that hit nothing
    plain = dict(spec)
    plain.pop('reflectivity')
    img_p = on.render(plain, return_aux=False)
    assert np.abs(img_p - img_n).max() > 0.1
    assert np.all(img_n[(hit_n < 0).all(0)] == 0)


@pytest.mark.parametrize('general', [False, True])
def test_mirror_closed_form_equals_autograd(general):
    """The closed-form reverse pass through the bounce (secondary object; P, r, n_w chain into the
    primary object's transform) == float64 autograd of the same forward with constant masks."""
    spec = scenes.mirror_scene(n=40, general=general)
    ps = oc.PackedScene.from_spec(spec, camera_grad=0)
    img, hit, hit2 = oc.render_forward_secondary(ps)
    dl = np.random.RandomState(1).normal(size=img[0].shape).astype(np.float32)
    oc.lib().orc_set_f64_record(1)
    try:
        grad = oc.render_backward(ps, dl, hit)
    finally:
        oc.lib().orc_set_f64_record(0)
    dlt = torch.from_numpy(dl).double()
    _, im, g = og.gradients(spec, hit[0], lambda im_: (im_ * dlt).sum(), hit2_index=hit2[0])
    assert np.abs(im - img[0]).max() < 1e-4
    gc = oc.split_grad(grad[0], ps.N)
    for k in ('w2o', 'material', 'light_dir', 'light_int', 'look_at'):
        ref, a = g[k], gc[k]
        if k == 'w2o':
            ref = ref[:, :3, :]
        if k == 'material':
            ref, a = np.delete(ref, 3, 1), np.delete(a, 3, 1)
        assert np.max(np.abs(a - ref)) <= 1e-10 * np.max(np.abs(ref)), k


def test_mirror_autograd_finite_differences():
    spec = scenes.mirror_scene(n=32)
    ps = oc.PackedScene.from_spec(spec, camera_grad=0)
    _, hit, hit2 = oc.render_forward_secondary(ps)
    w = torch.from_numpy(np.random.RandomState(3).normal(size=(32, 32, 3)))
    loss_fn = lambda im: (im * w).sum()
    _, _, g = og.gradients(spec, hit[0], loss_fn, hit2_index=hit2[0])

    def f(name, index, eps):
        p = og.leaf_params(spec)
        with torch.no_grad():
            p[name][index] += eps
            return float(loss_fn(og.forward(spec, p, hit[0], hit2_index=hit2[0])))
    for name, index in (('w2o', (0, 0, 3)), ('w2o', (1, 1, 3)), ('w2o', (1, 0, 0)), ('w2o', (2, 2, 3)), ('w2o', (2, 0, 2)),
                        ('w2o', (3, 1, 3)), ('material', (3, 1)), ('material', (0, 5)), ('light_dir', (0,)), ('look_at', (1,))):
        eps = 1e-8 if name == 'w2o' else 1e-6
        fd = (f(name, index, eps) - f(name, index, -eps)) / (2 * eps)
        assert abs(fd - g[name][index]) <= 2e-4 * max(1.0, abs(fd)), (name, index, fd, g[name][index])
