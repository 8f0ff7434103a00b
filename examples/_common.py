"""Shared pieces of the example ports: the two materials and the centre tensors every root
script of the reference starts from (optimize_brightness.py:19-27, match_mirror.py:16-24),
and the frame-dumping training loop they all end with."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reversible_raytracer_b200.scene import Material  # noqa: E402

GREEN = ((0.2, 0.9, 0.4), 0.3, 0.7, 0.5, 50.)        # Material(color, ks, kd, ka, shininess)
PINK = ((0.87, 0.1, 0.507), 0.3, 0.9, 0.4, 50.)


def materials():
    return Material(*GREEN), Material(*PINK)


def centres(device='cuda'):
    return [torch.tensor(c, device=device) for c in ([-.5, -.5, 4.], [.5, .5, 4.])]


def run(train, steps, frame, out=None, writer=None):
    """`steps` optimiser steps; after each one the current frame goes to out/<i>.png."""
    losses = []
    for i in range(1, steps + 1):
        losses.append(train())
        print('Step', i, losses[-1])
        if out is not None:
            writer(os.path.join(out, '%d.png' % i), frame())
    return losses
