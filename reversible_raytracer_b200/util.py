"""Helpers mirrored from the reference's util.py that sit on the render path.

Only `broadcasted_switch` (util.py:27-28) is on the hot path; `get_epsilon`
(util.py:23-24) is the optimiser loops' learning-rate decay.  Image I/O
(draw / drawWithMarkers, util.py:44-55) is kept as a thin NumPy/PIL writer so the
example scripts can dump frames; matplotlib is not required.
"""
import numpy as np
import torch


def broadcasted_switch(a, b, c):
    """util.py:27-28: T.switch(a.dimshuffle(0, 1, 'x'), b, c)"""
    b = torch.as_tensor(b, dtype=torch.float32, device=a.device) if not isinstance(b, torch.Tensor) else b
    c = torch.as_tensor(c, dtype=torch.float32, device=a.device) if not isinstance(c, torch.Tensor) else c
    return torch.where(a.bool().unsqueeze(-1), b, c)


def transNorm(transM, vec):
    """util.py:31-42: vec[x, y, :] @ transM[:3, :3] (normals through the transposed matrix); dense helper."""
    m = torch.as_tensor(transM, dtype=vec.dtype, device=vec.device)[:3, :3]
    return torch.tensordot(vec[:, :, :3], m, dims=([2], [0]))


def initialize_weight(n_vis, n_hid, W_name=None, numpy_rng=None, rng_dist='uniform', device=None):
    """util.py:10-20: Glorot-uniform (or 0.01 * normal) [n_vis, n_hid] float32 weight as a trainable tensor
    (the reference returns a theano.shared named W_name)."""
    rng = numpy_rng if numpy_rng is not None else np.random
    if 'uniform' in rng_dist:
        b = np.sqrt(6. / (n_vis + n_hid))
        W = rng.uniform(low=-b, high=b, size=(n_vis, n_hid))
    elif rng_dist == 'normal':
        W = 0.01 * rng.normal(size=(n_vis, n_hid))
    else:
        raise ValueError('rng_dist must be "uniform" or "normal"')
    from .transform import default_device
    t = torch.tensor(np.asarray(W, dtype=np.float32), device=device if device is not None else default_device())
    return t.requires_grad_(True)


def get_epsilon(epsilon, n, i):
    """Decaying learning rate, util.py:23-24."""
    return float(epsilon / (1 + i / float(n)))


def draw(fname, im):
    """util.py:54-55 (scipy.misc.imsave): min-max scaled 8-bit image."""
    from PIL import Image
    a = im.detach().cpu().numpy() if isinstance(im, torch.Tensor) else np.asarray(im)
    a = a.astype(np.float64)
    lo, hi = float(a.min()), float(a.max())
    a = (a - lo) / (hi - lo) if hi > lo else np.zeros_like(a)
    Image.fromarray((a * 255).astype(np.uint8)).save(fname)


def drawWithMarkers(fname, im):
    """util.py:44-52: the two marker rectangles of optimize_brightness.py."""
    a = im.detach().cpu().numpy().copy() if isinstance(im, torch.Tensor) else np.array(im, dtype=np.float64)
    for (cx, cy) in ((85, 90), (90, 50)):
        x0, x1, y0, y1 = cx - 3, cx + 3, cy - 3, cy + 3
        if a.shape[0] > y1 and a.shape[1] > x1:
            a[y0:y1 + 1, [x0, x1]] = (1, 0, 0)
            a[[y0, y1], x0:x1 + 1] = (1, 0, 0)
    draw(fname, a)
