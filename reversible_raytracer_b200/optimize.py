"""Optimiser loops -- thin PyTorch re-host of the reference's optimize.py, the
direct CALLER of the hot path (it owns T.grad + the SGD update, optimize.py:19-29).

The reference compiles `loss` (a symbolic expression) into one Theano function;
here `loss` is a CLOSURE that re-renders with the current parameter values and
returns a scalar tensor, e.g.

    train = GDOptimizer().optimize([center1, center2],
                                   lambda: -scene.build()[90, 85].sum() - scene.build()[50, 90].sum())
    for i in range(90): print(train(0.0008))

Both call shapes found in the reference are accepted, told apart by the number of positional
arguments exactly as the reference's two signatures would bind them:
    optimize(tVars, loss[, momentum]) -> fn(lr)            optimize.py:19-29 (HEAD; momentum is
                                                           accepted and unused there, too)
    optimize(tVars, loss, lr, momentum) -> fn()            optimize_brightness.py:50-52,
                                                           match_mirror.py:48 (stale form)
`lr=` / `momentum=` may also be given by keyword.

CUDA-graph capture (graph='auto', the default) -- what it freezes.  The reference's compiled
function is symbolic; a captured closure is a RECORDING: only device-side state is live on replay.
Anything the closure reads on the HOST per call is frozen at capture time: Python scalars, a
changing `seed=` / `jitter=`, a data index, `setTransform`-style scene mutations, a new Material.
Guards: (1) right after capture the graph is replayed once with lr = 0 and its loss is compared
with an eager evaluation -- a mismatch drops the graph with a warning and stepping stays eager;
(2) `train.recapture()` invalidates the graph after you changed host-side state on purpose;
(3) `graph=False` never captures.
"""
import warnings

import torch

from .util import get_epsilon  # noqa: F401


class _WholeStep(object):
    """One kernel launch per optimise step (rrt_small_step_mse) for a `Scene.mse_cost` closure
    whose variables are exactly the parameters of the shapes' transform chains.  Raises
    ValueError when the scene does not qualify (the caller then takes the general path)."""

    def __init__(self, spec, tVars):
        import ctypes as C
        from . import _native as nat, render as R
        from .transform import as_tensor
        scene = spec['scene']
        if not torch.cuda.is_available():
            raise ValueError('no CUDA device')
        dev = scene.device()
        if dev.type != 'cuda':
            raise ValueError('scene is not on a CUDA device')
        st = scene._static(dev)
        prog, cam_prog = st['prog'], st['cam_prog']
        if prog is None or not prog.dynamic or cam_prog is None or cam_prog.dynamic or scene.camera.look_at.requires_grad:
            raise ValueError('needs chain-expressible shape transforms with parameters and a constant camera')
        if st['mat'] is None or st['light_t'] is None:
            raise ValueError('materials and light must be constants')
        if st['refl'] is not None:
            raise ValueError('the mirror bounce is not supported by the whole-step kernel')
        if {id(p) for p in prog.param_tensors} != {id(v) for v in tVars}:
            raise ValueError('the optimised variables must be exactly the parameters of the shape transforms')
        cfg = scene.config(spec['antialias_samples'], cull=False)
        if spec.get('linear'):                      # Scene.linear_cost: RRT_FLAG_LINEAR_COST
            from dataclasses import replace
            cfg = replace(cfg, linear_cost=1)
        N, S = len(scene.shapes), cfg.samples
        if N < 1 or N > 32 or S > 32 or (S & (S - 1)) or cfg.shadows or cfg.deterministic or cfg.n * cfg.n * S > (16 << 20):
            raise ValueError('not a small scene')
        if any(p.dtype != torch.float32 or p.device != dev for p in prog.param_tensors):
            raise ValueError('parameters must be float32 tensors on the scene device')
        # every check happens BEFORE the user's tensors are touched
        values = prog.values().detach().clone().contiguous()
        off = int(prog.const_block.numel())
        self.param_begin = off
        if off + sum(p.numel() for p in prog.param_tensors) != values.numel():
            raise ValueError('unexpected chain value layout')
        obj_type, _, mat, light, cam = scene.pack(dev)
        jit = scene._jitter_for(cfg.n, cfg.samples, spec['jitter'], spec['seed'], dev)
        self.target = as_tensor(spec['target']).to(dev, torch.float32).contiguous()
        if self.target.numel() != cfg.n * cfg.n * 3:
            raise ValueError('target must be [n, n, 3]')
        # parameters move INTO the chain's value buffer; the user's tensors become views of it
        for p in prog.param_tensors:
            p.data = values[off:off + p.numel()].view(p.shape)
            off += p.numel()
        G = nat.grad_size(N)
        z = lambda n, dt: torch.zeros(n, dtype=dt, device=dev)
        self.keep = dict(values=values, w2o=z(N * 12, torch.float32).reshape(N, 12), grad=z(G, torch.float32),
                         g_values=z(values.numel(), torch.float32), loss_acc=z(1, torch.float64),
                         loss_out=z(1, torch.float32), ticket=z(1, torch.int32), prog=prog)
        k = self.keep
        self.tables = R._Tables(cfg, obj_type, k['w2o'], mat.detach(), light.detach(), cam.detach(), jit)
        s = nat.RrtStep()
        s.ops, s.chain_begin, s.values = prog.ops.data_ptr(), prog.chain_begin.data_ptr(), values.data_ptr()
        s.num_values, s.param_begin = int(values.numel()), int(self.param_begin)
        s.grad, s.g_values, s.loss_acc = k['grad'].data_ptr(), k['g_values'].data_ptr(), k['loss_acc'].data_ptr()
        # the step's loss is written by the kernel straight into pinned (mapped) host memory: no copy node
        self.host = torch.zeros(1, dtype=torch.float32).pin_memory()
        s.loss_out, s.ticket = self.host.data_ptr(), k['ticket'].data_ptr()
        self.step = s
        cw = spec['channel_weight']
        self.cw = (C.c_float * 3)(*[float(v) for v in cw]) if cw is not None else None
        self.device, self.C, self.nat = dev, C, nat
        if dev.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self._launch = nat.lib().rrt_small_step_mse
        self._desc, self._step, self._target = C.byref(self.tables.desc), C.byref(self.step), self.target.data_ptr()
        self._host_np = self.host.numpy()

    def __call__(self, lr):
        # the host side of a one-launch step is a handful of microseconds of Python around ~15 us of GPU work,
        # so it is kept lean: no device context switch when the device is current already, the pinned loss is
        # read through a NumPy view instead of tensor indexing
        C, nat = self.C, self.nat
        self.step.lr = float(lr)
        dev = self.device
        if torch.cuda.current_device() == dev.index:
            stream = torch.cuda.current_stream(dev)
            rc = self._launch(self._desc, self._step, self._target, self.cw, None, C.c_void_p(stream.cuda_stream))
        else:
            with torch.cuda.device(dev):
                stream = torch.cuda.current_stream(dev)
                rc = self._launch(self._desc, self._step, self._target, self.cw, None, C.c_void_p(stream.cuda_stream))
        if rc:
            nat.check(rc, 'rrt_small_step_mse')
        stream.synchronize()
        return float(self._host_np[0])


class GDOptimizer(object):
    """Gradient descent: var <- var - lr * dloss/dvar (optimize.py:11-29).

    The reference compiles forward + T.grad + update into ONE Theano function; the
    equivalent here is a CUDA graph: after two eager warm-up calls the whole step (loss
    closure, torch.autograd.grad, in-place update) is captured once and every later
    train() is a single graph replay (graph='auto', the default; falls back to eager
    stepping if the closure cannot be captured, e.g. it synchronises).  graph=False keeps
    eager stepping."""

    def __init__(self):
        pass

    def optimize(self, tVars, loss, *args, graph='auto', **kw):
        if len(args) > 2 or set(kw) - {'lr', 'momentum'}:
            raise TypeError('optimize(tVars, loss[, momentum]) or optimize(tVars, loss, lr, momentum)')
        lr = kw.get('lr')
        if len(args) == 2:                         # stale 4-argument form: (lr, momentum)
            lr = args[0]
        # len(args) == 1 is HEAD's `momentum`, which the reference never uses either (optimize.py:19-29)
        if not callable(loss):
            raise TypeError('loss must be a callable returning a scalar tensor (eager re-host of the '
                            'symbolic loss expression of optimize.py:19)')
        tVars = list(tVars)
        for v in tVars:
            v.requires_grad_(True)
        default_lr = lr
        use_graph = bool(graph) and all(v.is_cuda for v in tVars)
        st = dict(calls=0, graph=None, value=None, lr=None, failed=False, whole_step=None)
        spec = getattr(loss, 'fused_spec', None)
        if spec is not None and use_graph:         # Scene.mse_cost: try the one-launch step
            try:
                st['whole_step'] = _WholeStep(spec, tVars)
            except ValueError as e:
                st['whole_step_refused'] = str(e)

        def step(step_lr):
            value = loss()
            grads = torch.autograd.grad(value, tVars, allow_unused=True)
            with torch.no_grad():
                for var, g in zip(tVars, grads):
                    if g is None:
                        continue
                    g = g.to(var.device)
                    if isinstance(step_lr, torch.Tensor):       # captured step: lr lives on the device
                        var.addcmul_(g, step_lr.to(g.dtype), value=-1.0)      # one kernel per variable
                    else:
                        var.sub_(g, alpha=float(step_lr))
            return value

        def side_stream():
            # Every step -- eager warm-up steps included -- runs on ONE private stream: autograd ties a
            # leaf's gradient accumulator to the stream it was created on, and a capture on another
            # stream would have to synchronise with it (cudaErrorStreamCaptureImplicit when that is
            # the legacy default stream and the closure's graph is still referenced, e.g. by
            # scene.shapes).
            if st.get('stream') is None:
                st['stream'] = torch.cuda.Stream(device=tVars[0].device)
            return st['stream']

        def capture():
            dev = tVars[0].device
            st['lr'] = torch.zeros((), dtype=torch.float32, device=dev)
            side = side_stream()
            side.wait_stream(torch.cuda.current_stream(dev))
            saved = [v.detach().clone() for v in tVars]
            with torch.cuda.stream(side):          # warm-up on the capture stream (PyTorch capture recipe)
                step(st['lr'])
            torch.cuda.current_stream(dev).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            st['host'] = torch.zeros((), dtype=torch.float32).pin_memory()
            with torch.cuda.graph(g, stream=side):   # its per-stream scratch (render._ticket) exists from the warm-up
                st['value'] = step(st['lr'])
                # the loss read-back is part of the graph: a copy node into pinned host memory
                st['host'].copy_(st['value'].detach().to(torch.float32), non_blocking=True)
            # validate the recording once: replay with lr = 0 (no update) against an eager evaluation
            g.replay()
            torch.cuda.current_stream(dev).synchronize()
            replayed = float(st['host'])
            with torch.no_grad():
                eager = float(loss().detach())
            with torch.no_grad():                  # lr was 0 during warm-up/capture; restore exactly anyway
                for v, s0 in zip(tVars, saved):
                    v.copy_(s0)
            if not abs(replayed - eager) <= 1e-3 * max(1.0, abs(eager)):
                raise RuntimeError('replayed loss %r != eager loss %r: the closure reads host-side state '
                                   'that a CUDA graph cannot see' % (replayed, eager))
            st['graph'] = g
            st['lr_value'] = None

        def train(step_lr=None):
            step_lr = default_lr if step_lr is None else step_lr
            if step_lr is None:
                raise TypeError('learning rate missing: call train(lr)')
            st['calls'] += 1
            if st['whole_step'] is not None:
                return st['whole_step'](step_lr)
            if use_graph and st['graph'] is None and not st['failed'] and st['calls'] > 2:
                try:
                    capture()
                except Exception as e:             # closure not capturable: keep stepping eagerly
                    warnings.warn('GDOptimizer: CUDA-graph capture of the step failed (%r); stepping eagerly' % (e,))
                    st['failed'] = True
                    st['graph'] = None
                    torch.cuda.synchronize()
            if st['graph'] is not None:
                if st.get('lr_value') != float(step_lr):       # the learning rate lives on the device: upload on change only
                    st['lr'].fill_(float(step_lr))
                    st['lr_value'] = float(step_lr)
                st['graph'].replay()
                torch.cuda.current_stream(st['lr'].device).synchronize()
                return float(st['host'])
            if use_graph:                          # eager steps of a to-be-captured optimiser: on the private stream
                cur, side = torch.cuda.current_stream(tVars[0].device), side_stream()
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    value = step(step_lr).detach()
                cur.wait_stream(side)
                return float(value)
            return float(step(step_lr).detach())

        def recapture():
            """Forget the captured graph (call after changing host-side state the loss closure
            reads: seeds, data indices, scene structure); the next train() captures again."""
            st['graph'], st['failed'], st['lr_value'] = None, False, None
            st['calls'] = max(st['calls'], 2)

        train.state = st
        train.recapture = recapture
        return train


class MGDAutoOptimizer(object):
    """Autoencoder trainer (optimize.py:63-84; orbit_experiments/optimize.py:68-97).

    `ae` must expose `params` (list of tensors) and `cost(X)` (root) or
    `cost(Xl, Xr)` (orbit: left/right camera views)."""

    def __init__(self, ae):
        self.ae = ae

    def optimize(self, train_data, lam=None, fixed_length=3, graph='auto'):
        """-> opt(lr) -> cost (root, optimize.py:68-84) or opt(i, lr) -> cost (orbit variant: sample index,
        orbit_experiments/optimize.py:68-97).  Like the reference's ONE compiled `train` function, the whole
        step (encoder, render, T.grad, update) becomes one CUDA-graph replay after two eager warm-up calls
        (graph='auto'; same guards as GDOptimizer: replay-vs-eager validation, eager fallback with a
        warning, `opt.recapture()`, graph=False).  The sample the orbit variant indexes is copied into a
        static buffer before every replay, so `i` stays live."""
        ae = self.ae
        for p in ae.params:
            p.requires_grad_(True)
        orbit = lam is not None or (hasattr(train_data, 'dim') and train_data.dim() >= 3)
        bias_scale = 0.1 if orbit else 1.0     # orbit variant: 1-D parameters (biases) move at 0.1 * lr
        #                                        (orbit_experiments/optimize.py:80-81)
        params = list(ae.params)
        tensor_data = isinstance(train_data, torch.Tensor)
        use_graph = bool(graph) and tensor_data and train_data.is_cuda and all(p.is_cuda for p in params)
        st = dict(calls=0, graph=None, failed=False, stream=None, lr=None, lr_value=None, host=None, sample=None)

        def cost_of(sample):
            return ae.cost(sample[0], sample[1]) if orbit else ae.cost(sample)

        def step(cost, lr):
            grads = torch.autograd.grad(cost, params, allow_unused=True)
            with torch.no_grad():
                for var, g in zip(params, grads):
                    if g is None:
                        continue
                    k = bias_scale if var.dim() == 1 else 1.0
                    if isinstance(lr, torch.Tensor):            # captured step: lr lives on the device
                        var.addcmul_(g, lr.to(g.dtype), value=-k)
                    else:
                        var.sub_(k * lr * g)
            return cost

        def side_stream():
            if st['stream'] is None:
                st['stream'] = torch.cuda.Stream(device=params[0].device)
            return st['stream']

        def capture(sample):
            dev = params[0].device
            st['lr'] = torch.zeros((), dtype=torch.float32, device=dev)
            st['sample'] = sample.detach().clone()               # static input of the recording
            side = side_stream()
            side.wait_stream(torch.cuda.current_stream(dev))
            saved = [v.detach().clone() for v in params]
            with torch.cuda.stream(side):
                step(cost_of(st['sample']), st['lr'])
            torch.cuda.current_stream(dev).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            st['host'] = torch.zeros((), dtype=torch.float32).pin_memory()
            with torch.cuda.graph(g, stream=side):
                value = step(cost_of(st['sample']), st['lr'])
                st['host'].copy_(value.detach().to(torch.float32), non_blocking=True)
            g.replay()                                           # lr = 0: validates without moving anything
            torch.cuda.current_stream(dev).synchronize()
            replayed = float(st['host'])
            with torch.no_grad():
                eager = float(cost_of(st['sample']).detach())
                for v, s0 in zip(params, saved):
                    v.copy_(s0)
            if not abs(replayed - eager) <= 1e-3 * max(1.0, abs(eager)):
                raise RuntimeError('replayed cost %r != eager cost %r: the cost reads host-side state that a '
                                   'CUDA graph cannot see (e.g. unseeded anti-alias jitter drawn per call)' % (replayed, eager))
            st['graph'] = g

        def run(sample, lr):
            st['calls'] += 1
            if use_graph and st['graph'] is None and not st['failed'] and st['calls'] > 2:
                try:
                    capture(sample)
                except Exception as e:     # noqa: BLE001
                    warnings.warn('MGDAutoOptimizer: CUDA-graph capture of the step failed (%r); stepping eagerly' % (e,))
                    st['failed'], st['graph'] = True, None
                    torch.cuda.synchronize()
            if st['graph'] is not None:
                if st['lr_value'] != float(lr):
                    st['lr'].fill_(float(lr))
                    st['lr_value'] = float(lr)
                if orbit:
                    st['sample'].copy_(sample)                   # the indexed sample stays live across replays
                st['graph'].replay()
                torch.cuda.current_stream(st['lr'].device).synchronize()
                return float(st['host'])
            if use_graph:                  # eager steps of a to-be-captured trainer: on the private stream
                cur, side = torch.cuda.current_stream(params[0].device), side_stream()
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    value = step(cost_of(sample), lr).detach()
                cur.wait_stream(side)
                return float(value)
            return float(step(cost_of(sample), lr).detach())

        if orbit:
            def opt(i, lr):
                return run(train_data[i], lr)
        else:
            def opt(lr):
                return run(train_data[0], lr)

        def recapture():
            st['graph'], st['failed'], st['lr_value'] = None, False, None
            st['calls'] = max(st['calls'], 2)
        opt.state, opt.recapture = st, recapture
        return opt

    def optimizeADAM(self, train_data, beta1=0.1, beta2=0.001, epsilon=1e-8, l=1e-8):
        """optimize.py:86-124 (root variant): the ADAM form of the autoencoder trainer, with the
        reference's own constants and its 5x step on 1-D parameters.  Returns
        (opt(lr) -> cost, get_grad, get_gradb) like the reference (gradients of the last
        capsule's weight and bias)."""
        ae = self.ae
        for p in ae.params:
            p.requires_grad_(True)
        m = [torch.zeros_like(p) for p in ae.params]
        v = [torch.zeros_like(p) for p in ae.params]
        state = dict(t=1.0)

        def grads_now():
            return torch.autograd.grad(ae.cost(train_data[0]), ae.params, allow_unused=True)

        def opt(lr):
            cost = ae.cost(train_data[0])
            grads = torch.autograd.grad(cost, ae.params, allow_unused=True)
            t = state['t']
            with torch.no_grad():
                for p, g, m_, v_ in zip(ae.params, grads, m, v):
                    if g is None:
                        continue
                    b1_t = 1 - (1 - beta1) * (l ** (t - 1))
                    m_.copy_(b1_t * g + (1 - b1_t) * m_)
                    v_.copy_(beta2 * g * g + (1 - beta2) * v_)
                    m_hat = m_ / (1 - (1 - beta1) ** t)
                    v_hat = v_ / (1 - (1 - beta2) ** t)
                    p.sub_((5.0 if p.dim() == 1 else 1.0) * lr * m_hat / (torch.sqrt(v_hat) + epsilon))
            state['t'] = t + 1
            return float(cost.detach())

        return opt, (lambda: grads_now()[-2]), (lambda: grads_now()[-1])
