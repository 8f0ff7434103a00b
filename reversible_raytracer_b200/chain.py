"""Native parameter -> matrix chain.

The reference evaluates translate / scale / rotate / `*` / `.inverse()`
(transform.py:32-38, 56-122) symbolically inside the compiled Theano function, and
`T.grad` differentiates through them.  In eager PyTorch that is a few hundred tiny
kernels per optimisation step.  Here the STRUCTURE of every shape's transform is
compiled once into a small op table, and each step runs

    values = cat(constants, live parameter tensors)          (1 torch kernel, autograd-aware)
    tables = rrt_chain_forward(ops, values)                   (1 kernel: every w2o / camera matrix)
    ... render ...
    g_values = rrt_chain_backward(ops, values, dL/dtables)    (1 kernel)

torch autograd then hands g_values back to the user's tensors through `cat`.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as nat


def flatten(transform, inverted=False):
    """Transform expression -> list of (kind, inverted, args) whose left-to-right matrix
    product equals transform.m (or .mInv when inverted).  None if not expressible."""
    e = transform._expr
    if e is None:
        return None
    tag = e[0]
    if tag == 'I':
        return []
    if tag in ('T', 'S'):
        return [(tag, inverted, (e[1],), e[2])]
    if tag == 'R':
        return [('R', inverted, (e[1], e[2]), e[3])]
    if tag == 'inv':
        return flatten(e[1], not inverted)
    if tag == 'mul':
        a, b = flatten(e[1], inverted), flatten(e[2], inverted)
        if a is None or b is None:
            return None
        return (b + a) if inverted else (a + b)       # (A.B)^-1 = B^-1 . A^-1
    return None


class ChainProgram(object):
    """Compiled op table for a list of Transforms (one output row of 12 floats each)."""

    def __init__(self, transforms, device):
        chains = [flatten(t) for t in transforms]
        if any(c is None for c in chains) or any(len(c) > nat.CHAIN_MAX_OPS for c in chains):
            raise ValueError('transform not expressible as a chain of translate/scale/rotate')
        self.device = device
        self.num_chains = len(chains)
        consts, self.param_tensors, slot_of = [], [], {}
        n_const = 0
        # first pass: constants get fixed offsets in the constant block
        plan = []
        for chain in chains:
            row = []
            for kind, inv, args, is_param in chain:
                offs = []
                for a in args:
                    if is_param:
                        offs.append(('p', a))
                    else:
                        v = a.detach().reshape(-1).to(torch.float32).cpu().numpy()
                        offs.append(('c', n_const))
                        consts.append(v)
                        n_const += v.size
                row.append((kind, inv, offs))
            plan.append(row)
        # parameters follow the constant block, each distinct tensor once
        p_off = n_const
        for row in plan:
            for kind, inv, offs in row:
                for j, (tag, a) in enumerate(offs):
                    if tag == 'p':
                        key = id(a)
                        if key not in slot_of:
                            slot_of[key] = p_off
                            self.param_tensors.append(a)
                            p_off += a.numel()
                        offs[j] = ('c', slot_of[key])
        self.num_values = max(p_off, 1)
        kind_id = {'T': nat.CHAIN_TRANSLATE, 'S': nat.CHAIN_SCALE, 'R': nat.CHAIN_ROTATE}
        ops, begin = [], [0]
        for row in plan:
            for kind, inv, offs in row:
                ops.append([kind_id[kind] | (nat.CHAIN_INVERT if inv else 0), offs[0][1],
                            offs[1][1] if len(offs) > 1 else 0, 0])
            begin.append(len(ops))
        if not ops:                                # identity-only chains: keep the table non-empty (non-NULL)
            ops = [[0, 0, 0, 0]]
        self.ops = torch.tensor(np.asarray(ops, dtype=np.int32).reshape(-1, 4), device=device)
        self.chain_begin = torch.tensor(np.asarray(begin, dtype=np.int32), device=device)
        self.const_block = torch.tensor(np.concatenate(consts) if consts else np.zeros(0, dtype=np.float32),
                                        dtype=torch.float32, device=device)
        self.dynamic = len(self.param_tensors) > 0
        self._static_out = None

    def values(self):
        parts = [self.const_block] + [p.reshape(-1).to(self.device, torch.float32) for p in self.param_tensors]
        v = torch.cat(parts) if len(parts) > 1 else self.const_block
        if v.numel() == 0:
            v = torch.zeros(1, dtype=torch.float32, device=self.device)
        return v

    def evaluate(self):
        """-> [num_chains, 12] float32, differentiable w.r.t. the live parameter tensors."""
        if not self.dynamic:
            if self._static_out is None:
                self._static_out = _ChainFn.apply(self.values(), self)
            return self._static_out
        return _ChainFn.apply(self.values(), self)


class _ChainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, values, prog):
        values = values.contiguous()
        out = torch.empty((prog.num_chains, 12), dtype=torch.float32, device=values.device)
        with torch.cuda.device(values.device):
            rc = nat.lib().rrt_chain_forward(prog.ops.data_ptr(), prog.chain_begin.data_ptr(), prog.num_chains,
                                             values.data_ptr(), out.data_ptr(),
                                             C.c_void_p(torch.cuda.current_stream(values.device).cuda_stream))
        nat.check(rc, 'rrt_chain_forward')
        ctx.prog = prog
        ctx.save_for_backward(values)
        return out

    @staticmethod
    def backward(ctx, g_out):
        (values,) = ctx.saved_tensors
        prog = ctx.prog
        g_out = g_out.contiguous()
        g_values = torch.empty_like(values)
        with torch.cuda.device(values.device):
            rc = nat.lib().rrt_chain_backward(prog.ops.data_ptr(), prog.chain_begin.data_ptr(), prog.num_chains,
                                              values.data_ptr(), g_out.data_ptr(), g_values.data_ptr(),
                                              int(values.numel()),
                                              C.c_void_p(torch.cuda.current_stream(values.device).cuda_stream))
        nat.check(rc, 'rrt_chain_backward')
        return g_values, None
