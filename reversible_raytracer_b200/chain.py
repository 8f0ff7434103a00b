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
    product equals transform.m (or .mInv when inverted); args are transform._Arg objects.
    None if not expressible."""
    e = transform._expr
    if e is None:
        return None
    tag = e[0]
    if tag == 'I':
        return []
    if tag in ('T', 'S'):
        return [(tag, inverted, (e[1],))]
    if tag == 'R':
        return [('R', inverted, (e[1], e[2]))]
    if tag == 'inv':
        return flatten(e[1], not inverted)
    if tag == 'mul':
        a, b = flatten(e[1], inverted), flatten(e[2], inverted)
        if a is None or b is None:
            return None
        return (b + a) if inverted else (a + b)       # (A.B)^-1 = B^-1 . A^-1
    return None


class _Structure(object):
    """Device-side op table of a list of chains.  Depends only on the STRUCTURE of the
    expressions (op kinds, constant values, which live parameters are shared between ops),
    not on the identity of the parameter tensors, so it is cached: a loss closure that rebuilds
    its shapes on every call (the reference's decoders do) compiles -- and uploads -- it once."""

    def __init__(self, key, device):
        chains, n_const, n_values = key
        kind_id = {'T': nat.CHAIN_TRANSLATE, 'S': nat.CHAIN_SCALE, 'R': nat.CHAIN_ROTATE}
        ops, begin, consts = [], [0], np.zeros(n_const, dtype=np.float32)
        for chain in chains:
            for kind, inv, args in chain:
                offs = []
                for tag, off, payload in args:
                    offs.append(off)
                    if tag == 'c':
                        v = np.frombuffer(payload, dtype=np.float32)
                        consts[off:off + v.size] = v
                ops.append([kind_id[kind] | (nat.CHAIN_INVERT if inv else 0), offs[0], offs[1] if len(offs) > 1 else 0, 0])
            begin.append(len(ops))
        if not ops:                                # identity-only chains: keep the table non-empty (non-NULL)
            ops = [[0, 0, 0, 0]]
        self.ops = torch.tensor(np.asarray(ops, dtype=np.int32).reshape(-1, 4), device=device)
        self.chain_begin = torch.tensor(np.asarray(begin, dtype=np.int32), device=device)
        self.const_block = torch.tensor(consts, dtype=torch.float32, device=device)
        self.num_values = max(n_values, 1)
        self.zero1 = torch.zeros(1, dtype=torch.float32, device=device)


_STRUCTURES = {}
_STRUCTURES_MAX = 256


class _LeafRef(object):
    """a whole leaf tensor as one parameter slot (see ChainProgram)"""
    __slots__ = ('src', 'key')

    def __init__(self, base):
        self.src, self.key = base, ('leaf', id(base))


def _whole_leaf(arg):
    """(leaf, element offset of the view inside it) when `arg` is a contiguous view of a contiguous leaf"""
    if arg._view is None:
        return None
    base, size, stride, offset = arg._view
    n = 1
    for d, st in zip(reversed(size), reversed(stride)):     # contiguous: strides of a dense row-major block
        if d != 1 and st != n:
            return None
        n *= d
    rel = offset - base.storage_offset()
    if not base.is_contiguous() or rel < 0 or rel + n > base.numel():
        return None
    return base, rel


class ChainProgram(object):
    """Compiled op table for a list of Transforms (one output row of 12 floats each), bound to the
    CURRENT live parameter tensors of those transforms."""

    def __init__(self, transforms, device):
        chains = [flatten(t) for t in transforms]
        if any(c is None for c in chains) or any(len(c) > nat.CHAIN_MAX_OPS for c in chains):
            raise ValueError('transform not expressible as a chain of translate/scale/rotate')
        self.device = device
        self.num_chains = len(chains)
        # pass 1: constants get fixed offsets in the constant block (in order of appearance)
        n_const = 0
        plan = []
        for chain in chains:
            row = []
            for kind, inv, args in chain:
                offs = []
                for a in args:
                    if a.param:
                        offs.append(['p', a, None])
                    else:
                        offs.append(['c', n_const, a.host.tobytes()])
                        n_const += a.host.size
                row.append((kind, inv, offs))
            plan.append(row)
        # pass 2: live parameters follow the constant block, each distinct tensor once
        # (the _Arg objects are kept, not their tensors: an argument that is a view of a leaf is taken
        # again at every evaluation, see transform._Arg).  A CONTIGUOUS view of a contiguous leaf --
        # `translate(p[:3]) * scale(p[3:])`, test_balls.py:27 -- does not even need that: the whole leaf
        # gets one slot and the op reads its argument at (slot + offset of the view), so the value vector
        # is cat(constants, leaves), autograd goes straight to the leaves, and GDOptimizer's one-launch
        # step recognises the leaves as the chains' parameters.
        self.param_args, slot_of, p_off = [], {}, n_const
        for row in plan:
            for kind, inv, offs in row:
                for o in offs:
                    if o[0] == 'p':
                        arg = o[1]
                        numel = arg.src.numel()
                        leaf = _whole_leaf(arg)
                        if leaf is not None:
                            base, rel = leaf
                            bkey = ('leaf', id(base))
                            if bkey not in slot_of:
                                slot_of[bkey] = p_off
                                self.param_args.append(_LeafRef(base))
                                p_off += base.numel()
                            o[1], o[2] = slot_of[bkey] + rel, numel
                            continue
                        if arg.key not in slot_of:
                            slot_of[arg.key] = p_off
                            self.param_args.append(arg)
                            p_off += numel
                        o[1], o[2] = slot_of[arg.key], numel
        key = (tuple(tuple((kind, inv, tuple(tuple(o) for o in offs)) for kind, inv, offs in row) for row in plan),
               n_const, p_off)
        st = _STRUCTURES.get((key, str(device)))
        if st is None:
            if len(_STRUCTURES) >= _STRUCTURES_MAX:
                _STRUCTURES.clear()
            st = _STRUCTURES[(key, str(device))] = _Structure(key, device)
        self.structure = st
        self.ops, self.chain_begin, self.const_block, self.num_values = st.ops, st.chain_begin, st.const_block, st.num_values
        self.dynamic = len(self.param_args) > 0
        self._static_out = None

    @property
    def param_tensors(self):
        """the live parameter tensors, in slot order (views of leaves are re-taken on every access)"""
        return [a.src for a in self.param_args]

    def values(self):
        """cat(constants, live parameters) as float32 on the program's device (the cast is done
        here, at evaluation time, so in-place updates of non-float32 parameters are seen)."""
        parts = [self.const_block] + [p.reshape(-1).to(self.device, torch.float32) for p in self.param_tensors]
        v = torch.cat(parts) if len(parts) > 1 else self.const_block
        if v.numel() == 0:
            v = self.structure.zero1
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
