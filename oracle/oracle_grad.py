"""oracle_grad -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Gradient oracle: float64 torch-CPU restatement of the forward pass of
oracle_numpy.render with every mask treated as a CONSTANT, differentiated by
torch.autograd.  This is the semantics `T.grad(loss, params)` has on the
reference graph (/root/reference/optimize.py:25, :73;
orbit_experiments/optimize.py:76) under FAST_RUN: a ray contributes only through
the one shape that won its depth test, only if it hit, and through `clip` only on
the closed interval [0,1].

PARITY UNPINNED by the reference: it ships no gradient values, no gradient test
and no loss curve.  What pins this file instead: (1) its forward equals
oracle_numpy.render (tested), which is pinned to the reference's golden renders;
(2) central finite differences of that forward on mask-stable scenes (tested).

The hit masks are inputs (`hit_index`, int32[S,n,n], -1 = background) so that the
oracle differentiates exactly the winner selection the float32 path made.
"""
import numpy as np
import torch

from . import oracle_numpy as on

F64 = torch.float64


def effective_rays(spec):
    """Per IMAGE pixel camera-space ray r_eff[s,a,b,:] (float32 values, returned
    as float64 tensor): rays[b,a] for the root variant (one spatial transpose,
    transform.py:46 applied once by shape.w2o), rays[a,b] for the orbit variant
    (applied twice: camera.o2w then shape.w2o)."""
    n, S = int(spec['n']), int(spec['samples'])
    out = np.zeros((S, n, n, 3), dtype=np.float32)
    for s in range(S):
        sdx = (spec['jitter_x'][:, :, s] + np.float32(s)) / np.float32(S)
        sdy = (spec['jitter_y'][:, :, s] + np.float32(s)) / np.float32(S)
        _, rays = on.make_rays(n, n, sdx, sdy)
        out[s] = rays.transpose(1, 0, 2) if spec.get('cam_o2w') is None else rays
    return torch.from_numpy(out).to(F64)


def leaf_params(spec):
    """float64 leaf tensors for everything the reverse pass can reach."""
    p = dict(
        w2o=torch.tensor(np.asarray(spec['w2o'], dtype=np.float64)),
        material=torch.tensor(np.asarray(spec['material'], dtype=np.float64)),
        light_dir=torch.tensor(np.asarray(spec['light_dir'], dtype=np.float64)),
        light_int=torch.tensor(np.asarray(spec['light_int'], dtype=np.float64)),
        look_at=torch.tensor(np.asarray(spec['look_at'], dtype=np.float64)),
    )
    cam = spec.get('cam_o2w')
    p['cam_o2w'] = torch.tensor(np.eye(4) if cam is None else np.asarray(cam, dtype=np.float64))
    for v in p.values():
        v.requires_grad_(True)
    return p


def _hit_and_shade(spec, params, k, o, d, Lh, clip_closed):
    """One object k, object-space ray origin o [3] (or [P,3]) and directions d [P,3]:
    -> (t [P], object-space normal [P,3], rgb [P,3]); masks are the caller's business."""
    mat = params['material'][k]
    shader = spec['shader']
    ob = o if o.dim() == 2 else o[None, :].expand_as(d)
    if spec['obj_type'][k] == on.SPHERE:                # shape.py:109-138
        pd = (d * ob).sum(1)
        vn = (d * d).sum(1)
        det = pd * pd - vn * ((ob * ob).sum(1) - 1.0)
        t = (-pd - torch.sqrt(det)) / vn
        p = ob + t[:, None] * d
        nrm = p / torch.sqrt((p * p).sum(1))[:, None]
    else:                                               # shape.py:25-69
        t = -ob[:, 2] / d[:, 2]
        nrm = torch.zeros_like(d)
        nrm[:, 2] = torch.where(ob[:, 2].detach() > 0, 1.0, -1.0)
    if shader == 'depth':                               # shader.py:14-20
        rgb = (1.0 - t / float(spec['max_depth']))[:, None] * torch.ones(3, dtype=F64)
    else:                                               # shader.py:28-53
        # shininess is a constant in every reference script (no d/d shininess is
        # ever requested; it would be NaN for rv<0): detached here.
        ka, kd, ks, sh = mat[0], mat[1], mat[2], mat[3].detach()
        ndl = -(nrm @ Lh)
        ph = ka + kd * ndl
        if shader == 'phong':
            rm = 2.0 * ndl[:, None] * nrm + Lh
            rv = rm @ params['look_at']
            ph = ph + ks * torch.pow(rv, sh)
        col = ph[:, None] * mat[4:7][None, :] * params['light_int'][None, :]
        with torch.no_grad():
            inside = (col >= 0) & (col <= 1) if clip_closed else (col > 0) & (col < 1)
            const = torch.clamp(col, 0, 1)
        rgb = torch.where(inside, col, const)           # clip with constant mask
    return t, nrm, rgb


def forward(spec, params, hit_index, clip_closed=True, hit2_index=None):
    """Differentiable float64 image[n,n,3] given constant winners `hit_index` (and, for the mirror
    bounce -- spec['reflectivity'], RRT_FLAG_MIRROR in include/rrt_b200.h --, the constant secondary
    winners `hit2_index`)."""
    n, S = int(spec['n']), int(spec['samples'])
    R = effective_rays(spec)                                  # [S,n,n,3]
    cam = params['cam_o2w']
    Cm, ct = cam[:3, :3], cam[:3, 3]
    Lh = params['light_dir'] / torch.sqrt((params['light_dir'] ** 2).sum())   # scene.py:83-86
    hit_index = torch.as_tensor(np.asarray(hit_index))
    refl = spec.get('reflectivity')
    if refl is not None:
        hit2_index = torch.as_tensor(np.asarray(hit2_index))
    image = torch.zeros(n, n, 3, dtype=F64)
    for s in range(S):
        img_s = torch.zeros(n, n, 3, dtype=F64)
        for k in range(len(spec['obj_type'])):
            sel = (hit_index[s] == k).nonzero(as_tuple=True)
            if sel[0].numel() == 0:
                continue
            A, b = params['w2o'][k, :3, :3], params['w2o'][k, :3, 3]
            dw = R[s][sel] @ Cm.T                               # world dirs   [P,3]
            o = A @ ct + b                                      # object-space origin
            d = dw @ A.T                                        # object-space dirs
            t, nrm, rgb = _hit_and_shade(spec, params, k, o, d, Lh, clip_closed)
            if refl is not None:                                # one mirror bounce (extension)
                kr = float(refl[k])
                m = nrm @ A                                     # A^T n_o per row
                nw = m / torch.sqrt((m * m).sum(1))[:, None]
                dn = (dw * nw).sum(1)
                r = dw - 2.0 * dn[:, None] * nw
                P = ct[None, :] + t[:, None] * dw
                rgb2 = torch.zeros_like(rgb)
                j2s = hit2_index[s][sel]
                for j in range(len(spec['obj_type'])):
                    sub = (j2s == j).nonzero(as_tuple=True)[0]
                    if sub.numel() == 0:
                        continue
                    Aj, bj = params['w2o'][j, :3, :3], params['w2o'][j, :3, 3]
                    o2 = P[sub] @ Aj.T + bj[None, :]
                    d2 = r[sub] @ Aj.T
                    _, _, c2 = _hit_and_shade(spec, params, j, o2, d2, Lh, clip_closed)
                    rgb2 = rgb2.index_put((sub,), c2)
                rgb = (1.0 - kr) * rgb + kr * rgb2
            img_s = img_s.index_put(sel, rgb)
        image = image + img_s
    return image / S


def gradients(spec, hit_index, loss_fn, hit2_index=None):
    """Returns (loss float, image float64[n,n,3], grads dict of float64 arrays).
    loss_fn: image tensor -> scalar tensor."""
    params = leaf_params(spec)
    image = forward(spec, params, hit_index, hit2_index=hit2_index)
    loss = loss_fn(image)
    names = list(params.keys())
    gs = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
    grads = {k: (np.zeros(tuple(params[k].shape)) if g is None else g.numpy().copy())
             for k, g in zip(names, gs)}
    return float(loss.detach()), image.detach().numpy(), grads
