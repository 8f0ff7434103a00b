#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native differentiable ray tracer.

Metric (BASELINE.json): Mrays/s forward+backward on the synthetic stress scene C5
(4096 x 4096 image, 4 anti-alias samples, 1024 spheres), one "step" = one fused
forward + squared-error loss + reverse pass over the whole image; with N GPUs the
image's row slabs are sharded across ranks (total work fixed => strong scaling; contiguous
slabs of equal estimated COST, sharding.balanced_row_slabs, chosen during set-up) and the small
parameter-gradient vector + loss are summed by ONE kernel of our own over NVLink peer memory
(rrt_peer_allreduce; NCCL allreduce as the fallback).  The step's launches are recorded once
and replayed as one CUDA graph, the way an optimiser loop runs them (--no-graph-step: from Python).

    python bench.py [--gpus N] [--steps K] [--warmup W]            this framework
    python bench.py --impl reference [...]                         CPU restatement of the
        reference algorithm (oracle/oracle_c.c, all host threads) on a bounded sample
        of the same workload -- the reference itself is Python 2 + Theano and cannot
        run in this image.

Prints ONE JSON line on rank 0.  Every N reports `roofline`; N = 1 adds `cpu_baseline` and
`other_configs` (C1-C4 with CPU restatements beside them, C5g, culling); N > 1 adds
`exchange_check` and the sharded C4 decoder batch / training step.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = 'C5 synthetic stress: 4096x4096, S=4, 1024 spheres (translate*scale), fused fwd+mse+bwd'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--n', type=int, default=4096)
    ap.add_argument('--objects', type=int, default=1024)
    ap.add_argument('--samples', type=int, default=4)
    ap.add_argument('--general', action='store_true', help='C5g: rotated, non-uniformly scaled spheres')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the other BASELINE configs (C1/C3 step latency, C4, C5g)')
    ap.add_argument('--miss', action='store_true', help='diagnostic: move every sphere out of view (pure sweep, no hits)')
    ap.add_argument('--canonical', action='store_true',
                    help='headline with RRT_FLAG_CANONICAL_SWEEP (every pair in the reference arithmetic, no pre-filter)')
    ap.add_argument('--cpu-seconds', type=float, default=20.0)
    ap.add_argument('--workload', default='c5', choices=['c5', 'c4'],
                    help="c5 (default): the headline stress scene; c4: BASELINE config 4, the orbit autoencoder's "
                         "decoder batch (256 scenes x 2 views, 64x64, S=4), scene ranges sharded across ranks")
    ap.add_argument('--scenes', type=int, default=256)
    ap.add_argument('--no-graph-step', action='store_true',
                    help='launch the step\'s kernels from Python every time instead of replaying them as one CUDA graph')
    ap.add_argument('--uniform-slabs', action='store_true',
                    help='N > 1: row slabs of equal HEIGHT (default: contiguous slabs of equal estimated COST, from the '
                         'per-row hit counts of a first render; sharding.balanced_row_slabs)')
    return ap.parse_args()


# ------------------------------------------------------------------ CPU restatement legs
def cpu_fused_sample(args, rows, row_begin=None):
    """Times oracle_c's fused fwd+mse+bwd on `rows` rows of the workload (all host
    threads).  Together with cpu_small_configs the ONLY places bench.py executes anything under oracle/."""
    from oracle import oracle_c as oc
    from reversible_raytracer_b200 import workloads as W
    oc.use_all_cores()                      # torchrun sets OMP_NUM_THREADS=1
    tb = W.stress_tables(args.objects, general=args.general)
    n, S = args.n, args.samples
    rb = (n // 2 - rows // 2) if row_begin is None else row_begin
    ps = oc.PackedScene(n, S, tb['obj_type'], tb['w2o'], tb['material'], tb['light'], tb['camera'], tb['shader'],
                        tb['transpose'], seed=4321, row_begin=rb, row_count=rows)
    target = np.zeros((1, rows, n, 3), dtype=np.float32)
    t0 = time.perf_counter()
    oc.render_fused_mse(ps, target)
    dt = time.perf_counter() - t0
    return dt, rows * n * S, oc.num_threads()


def cpu_baseline(args, budget_s):
    dt, rays, threads = cpu_fused_sample(args, 8)              # calibrate (also warms the thread pool)
    dt, rays, threads = cpu_fused_sample(args, 8)
    rows = int(max(8, min(args.n, 8 * budget_s / max(dt, 1e-6))))
    dt, rays, threads = cpu_fused_sample(args, rows)
    return dict(value=rays / dt / 1e6, unit='Mrays/s', cores=threads, kind='port',
                sample='%d of %d rows of the same scene (n=%d, S=%d, N=%d), oracle_c fused fwd+mse+bwd, %.1f s, '
                       'gcc -O3 AVX2+FMA OpenMP' % (rows, args.n, args.n, args.samples, args.objects, dt))


def best_of(fn, k=3):
    fn()
    best = 1e30
    for _ in range(k):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_small_configs():
    """the reference's own configs C1-C4 at FULL size on the host CPU, restated
    (Theano cannot run here): CPU-A = oracle_numpy, the dense per-shape whole-image algorithm in the
    reference's structure (float32; forward), with its reverse pass as dense float64 torch-CPU autograd
    (oracle_grad, the closest stand-in for T.grad); CPU-B = oracle_c, per-pixel C loop, OpenMP on all host
    threads.  perf_counter, warm-up 1, best of 3.  One optimise step = forward + loss + reverse pass."""
    import torch
    from oracle import oracle_c as oc, oracle_numpy as on, oracle_grad as og, scenes
    oc.use_all_cores()
    threads = oc.num_threads()
    out = {}

    def one(name, spec, loss_fn, mse_target=None, views=1):
        ps = oc.PackedScene.from_spec(spec)
        rays = spec['n'] * spec['n'] * spec['samples']
        t_a_fwd = best_of(lambda: on.render(spec, return_aux=False))
        _, hit, _ = oc.render_forward(ps)
        t_a_step = best_of(lambda: og.gradients(spec, hit[0], loss_fn))
        t_b_fwd = best_of(lambda: oc.render_forward(ps, want_aux=False))
        if mse_target is not None:
            t_b_step = best_of(lambda: oc.render_fused_mse(ps, mse_target))
        else:
            dl = np.zeros((1, spec['n'], spec['n'], 3), dtype=np.float32)
            dl[0, 90 % spec['n'], 85 % spec['n']] = -1.0
            dl[0, 50 % spec['n'], 90 % spec['n']] = -1.0

            def step_b():
                oc.render_forward(ps)
                oc.render_backward(ps, dl)
            t_b_step = best_of(step_b)
        out[name] = dict(
            rays_per_step=rays * views,
            cpu_a_numpy=dict(forward_ms=round(t_a_fwd * 1e3 * views, 3), step_ms=round(t_a_step * 1e3 * views, 3),
                             Mrays_s_forward=round(rays / t_a_fwd / 1e6, 3), threads=torch.get_num_threads(),
                             what='forward = oracle_numpy.render (dense float32 whole-image passes per sample and shape, the '
                                  'reference\'s structure); step = oracle_grad forward + reverse (float64 torch-CPU autograd '
                                  'over the winners only -- sparser, i.e. cheaper, than Theano\'s dense T.grad graph)'),
            cpu_b_c=dict(forward_ms=round(t_b_fwd * 1e3 * views, 3), step_ms=round(t_b_step * 1e3 * views, 3),
                         Mrays_s_forward=round(rays / t_b_fwd / 1e6, 3), threads=threads,
                         what='oracle_c (per-pixel C, canonical order, OpenMP): forward; step = fused fwd+mse+bwd '
                              '(C1: forward + reverse pass of the two-pixel loss)'),
            extrapolated=(views > 1))

    sp = scenes.optimize_brightness()
    one('C1_optimize_brightness', sp, lambda im: -im[90, 85].sum() - im[50, 90].sum())
    sp = scenes.test_balls()
    tgt = np.zeros((1, 32, 32, 3), dtype=np.float32)
    one('C2_test_balls', sp, lambda im: ((im[:, :, 0]) ** 2).sum(), mse_target=tgt)
    sp = scenes.match_mirror()
    tgt = np.zeros((1, 128, 128, 3), dtype=np.float32)
    one('C3_match_mirror', sp, lambda im: (im ** 2).sum(), mse_target=tgt)
    sp = scenes.orbit((3.8307, -8.1441, 32), 0)
    tgt = np.zeros((1, 64, 64, 3), dtype=np.float32)
    one('C4_orbit_256x2', sp, lambda im: (im ** 2).sum(), mse_target=tgt, views=512)
    out['C4_orbit_256x2']['note'] = 'one 64x64 view timed, x512 views of the batch (the reference renders them one by one)'
    # BASELINE.md section 3, CPU-A on C5: the dense reference-structured algorithm cannot run the full size in
    # reasonable time (and Theano could not even compile it), so it is timed at n=192, N=64, S=4 and extrapolated
    # linearly in rays x objects, flagged as such
    sp = scenes.stress(n=192, num_objects=64, samples=4)
    t0 = time.perf_counter()
    on.render(sp, return_aux=False)
    t = time.perf_counter() - t0
    tests = 192.0 * 192 * 4 * 64
    full = 4096.0 * 4096 * 4 * 1024
    out['C5_stress'] = dict(cpu_a_numpy=dict(
        measured_at='n=192, N=64, S=4 (one forward, no warm-up)', forward_s=round(t, 3), ray_object_tests_per_s=round(tests / t, 1),
        extrapolated_full_size_forward_hours=round(t * full / tests / 3600.0, 2), threads=torch.get_num_threads(), extrapolated=True,
        what='oracle_numpy.render forward, extrapolated linearly in rays x objects to 4096^2 x 4 x 1024; CPU-B (oracle_c) on the '
             'REAL full-size scene is the `cpu_baseline` of this line'))
    return out


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # each step = a bounded sample sized from a calibration run to ~2 s
    dt, rays, threads = cpu_fused_sample(args, 8)
    dt, rays, threads = cpu_fused_sample(args, 8)
    rows = int(max(8, min(args.n, 8 * 2.0 / max(dt, 1e-6))))
    for _ in range(args.warmup):
        cpu_fused_sample(args, rows)
    t0 = time.perf_counter()
    total = 0
    for k in range(args.steps):
        _, rays, _ = cpu_fused_sample(args, rows)
        total += rays
    dt = time.perf_counter() - t0
    v = total / dt / 1e6
    sample = '%d of %d rows per step' % (rows, args.n)
    print(json.dumps(dict(
        impl='reference', metric='Mrays/s fwd+bwd', value=v, unit='Mrays/s', n_gpus=args.gpus, steps=args.steps,
        warmup=args.warmup, ms_per_step=dt / args.steps * 1e3, higher_is_better=True, scaling='strong',
        vs_baseline=None, dtype='f32', data='synthetic',
        config=dict(workload=WORKLOAD, n=args.n, samples=args.samples, objects=args.objects,
                    note='reference = CPU restatement (oracle/oracle_c.c); Theano/Python 2 cannot run here'),
        cpu_baseline=dict(value=v, unit='Mrays/s', cores=threads, kind='port', sample=sample),
        e2e=dict(value=v, unit='Mrays/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


# ------------------------------------------------------------------ clocks sampler
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonSwPowerCap: 'sw_power_cap',
            nv.nvmlClocksThrottleReasonHwSlowdown: 'hw_slowdown',
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: 'sw_thermal_slowdown',
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: 'hw_thermal_slowdown',
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: 'hw_power_brake',
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        med = float(np.median(self.samples)) if self.samples else None
        return dict(sm_mhz=med, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons))


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs next to its GPU BEFORE any pinned host buffer is touched
    (first-touch NUMA placement): with 8 ranks moving image slabs over PCIe every step, remote
    host memory is what the copies wait for.  Best effort; returns the CPU count or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        cpus = {i * 64 + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:       # noqa: BLE001
        pass
    return None


# ------------------------------------------------------------------ shared timing helper
class Timer(object):
    """K steps, each bracketed by CUDA events on the current stream, an L2 flush write between
    them (outside the timed intervals); barrier + synchronize on both sides; MAX over ranks."""

    def __init__(self, dev, world):
        import torch
        self.torch, self.dev, self.world = torch, dev, world
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def sync(self):
        import torch.distributed as dist
        self.torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, fn, steps, warm=1, flush=True):
        """-> (total ms over `steps` steps, max over ranks; per-step ms of this rank)"""
        import torch.distributed as dist
        torch = self.torch
        for _ in range(warm):
            fn()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.sync()
        for k in range(steps):
            if flush:
                self.flush.fill_(k & 0xff)
            evs[k][0].record()
            fn()
            evs[k][1].record()
        self.sync()
        per = [a.elapsed_time(b) for a, b in evs]
        t = torch.tensor([sum(per)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t), per


# ------------------------------------------------------------------ the other BASELINE configs (1 GPU)
def other_configs(dev, peak_tf, hbm_gbs, with_cpu=True):
    """Optimise-step latency of the reference's own small configs (C1 optimize_brightness.py,
    C3 match_mirror.py: 128x128, S=4) through the drop-in API + GDOptimizer (CUDA-graph
    replayed step incl. the device->host read of the loss), throughput of the batched
    autoencoder decoder workload C4 (256 scenes x 2 views, 64x64, S=4, fused fwd+mse+bwd) with
    both roofline fractions (it sits at the FP32/HBM ridge), C5g / C5 S=1 / culling
    for the record -- and the CPU restatements of C1-C4 beside them."""
    import torch
    from dataclasses import replace
    from reversible_raytracer_b200 import render as R, workloads as W, _native as nat
    from tools import latency as L
    out = {}
    train, _ = L.c1()
    out['C1_optimize_brightness_step_us'] = round(L.timeit(train, warm=6, iters=100), 1)
    train, _ = L.c1('graph')                    # the same loss as a weight image: fused kernel (RRT_FLAG_LINEAR_COST)
    out['C1_optimize_brightness_fused_step_us'] = round(L.timeit(train, warm=6, iters=100), 1)
    train, _ = L.c1(True)                       # Scene.linear_cost: whole step = one kernel launch
    if train.state.get('whole_step') is not None:
        out['C1_optimize_brightness_whole_step_kernel_us'] = round(L.timeit(train, warm=6, iters=100), 1)
    train, _ = L.c2(False)
    out['C2_test_balls_step_us'] = round(L.timeit(train, warm=6, iters=100), 1)
    train, _ = L.c2(True)
    out['C2_test_balls_fused_step_us'] = round(L.timeit(train, warm=6, iters=100), 1)
    train, _ = L.c2('whole')
    if train.state.get('whole_step') is not None:
        out['C2_test_balls_whole_step_kernel_us'] = round(L.timeit(train, warm=6, iters=100), 1)
    train, _ = L.c3(False)
    out['C3_match_mirror_step_us'] = round(L.timeit(train, warm=6, iters=100), 1)
    train, _ = L.c3('graph')
    out['C3_match_mirror_fused_step_us'] = round(L.timeit(train, warm=6, iters=100), 1)
    train, _ = L.c3(True)                       # Scene.mse_cost: whole step = one kernel launch (rrt_small_step_mse)
    if train.state.get('whole_step') is not None:
        out['C3_match_mirror_whole_step_kernel_us'] = round(L.timeit(train, warm=6, iters=100), 1)
    train, captured = L.c4_closure()            # the reference's closure style through Scene + GDOptimizer
    out['C4_one_scene_pair_through_the_API_step_us'] = dict(
        us=round(L.timeit(train, warm=6, iters=100), 1), cuda_graph=bool(captured()),
        note='orbit_experiments/test_optimization.py:17-44 style: materials, shapes, light, cameras and Scene rebuilt '
             'inside the loss closure on every call; 2 views of one scene, GDOptimizer step incl. loss read-back')

    def c4_entry(geom):
        fn, rays = L.c4(256, geom_grad_only=geom)
        # GPU time of the call alone: replayed from a CUDA graph (no Python between the launches)
        g, s = torch.cuda.CUDAGraph(), torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g, stream=s):
            fn()
        us_g = L.timeit(g.replay, warm=10, iters=200)
        us = L.timeit(fn, warm=5, iters=100)
        hits = L.c4_hits(256)
        flops = rays * 2 * 16.0 + hits * 320.0                   # 2 diagonal spheres + 320 per winning ray
        bytes_ = 2 * 256 * 64 * 64 * 3 * 4 * 1.0                # target read (the image is not requested by the trainer)
        return dict(us_per_batch=round(us, 1), us_per_batch_graph_replay=round(us_g, 1), Mrays_s=round(rays / us_g, 1),
                    roofline=dict(fp32_frac=round(flops / (us_g * 1e-6) / 1e12 / peak_tf, 4),
                                  hbm_frac=round(bytes_ / (us_g * 1e-6) / 1e9 / hbm_gbs, 4),
                                  algorithmic_flops=flops, algorithmic_bytes=bytes_, winning_rays=hits,
                                  note='latency / issue-bound: 8.4 M rays x ~70 useful flops; both fractions are '
                                       'reported because the workload sits at the FP32/HBM ridge (SURVEY.md 8d)'))
    out['C4_orbit_256x2_fused'] = c4_entry(1)
    out['C4_orbit_256x2_fused']['gradients'] = 'd/d w2o only (RRT_FLAG_NO_MATERIAL_GRAD: what every reference decoder needs)'
    out['C4_orbit_256x2_fused_all_gradients'] = c4_entry(0)

    def stress(general, fwd_only, samples=4, n=4096, N=1024, iters=3, cull=0, canonical=0):
        tb = W.stress_tables(N, general=general)
        t = lambda a: torch.from_numpy(a).to(dev)
        cfg = R.RenderConfig(n=n, samples=samples, shader=nat.SHADER_PHONG, transpose=1, seed=4321, cull=cull,
                             canonical_sweep=canonical)
        args = (t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']), t(tb['camera']))
        target, _, _ = R.render_forward(cfg, *args, None, want_hit=False)
        fn = (lambda: R.render_forward(cfg, *args, None, want_hit=False)) if fwd_only else \
             (lambda: R.render_fused_mse(cfg, *args, target, want_image=True))
        us = L.timeit(fn, warm=2, iters=iters)
        return dict(ms=round(us / 1e3, 3), Mrays_s=round(n * n * samples / us, 1))
    out['C5_S1_fused'] = stress(False, False, samples=1)
    rays = 4096.0 * 4096 * 4
    g = stress(True, False)
    g['frac_fp32_peak_28flop_per_test'] = round((rays * 1024 * 28) / (g['ms'] * 1e-3) / 1e12 / peak_tf, 4)
    gc = stress(True, False, canonical=1)
    gc['frac_fp32_peak_28flop_per_test'] = round((rays * 1024 * 28) / (gc['ms'] * 1e-3) / 1e12 / peak_tf, 4)
    g['canonical_sweep'] = gc
    g['note'] = ('general affine spheres: the pre-filter is the same 5-FMA quadratic form, so C5g costs what C5 costs; '
                 'credited flops are 28 per test (SURVEY.md 8d), executed FMA-lane flops 10 per test -- see roofline.executed')
    out['C5g_general_affine_fused'] = g
    # NOT roofline-accountable: conservative per-tile culling skips work (results bit-identical, tested)
    c = stress(False, False, cull=1, iters=5)
    c['note'] = 'RRT_FLAG_CULL: same bits as the exhaustive sweep, reported separately, never as a roofline fraction'
    out['C5_fused_with_culling'] = c
    if with_cpu:
        try:
            out['cpu_restatements_C1_C4'] = cpu_small_configs()
        except Exception as e:      # noqa: BLE001
            out['cpu_restatements_C1_C4'] = {'error': repr(e)}
    return out


# ------------------------------------------------------------------ C4 across GPUs (runs inside the default bench at N > 1)
def c4_sharded(dev, world, rank, timer, scenes=256):
    """BASELINE config 4 at N GPUs: (a) the decoder batch alone -- scene ranges per rank, no collective on
    the render path; (b) the whole autoencoder training step (examples/orbit_autoencoder.py): encoder
    forward/backward in stock PyTorch, the fused render launch, ONE flat NCCL allreduce of the 7.4 M
    encoder-weight gradients, SGD update.  Also with 256 scenes PER RANK (weak scaling):
    32 scenes per GPU are one ~25 us launch, i.e. pure launch latency."""
    import torch
    import importlib.util
    from reversible_raytracer_b200 import render as R, workloads as W, sharding as Sh
    out = {}
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def decoder(total_scenes):
        tb, tt = W.orbit_tables(total_scenes), W.orbit_tables(total_scenes, centre_noise=0.5)
        first, count = Sh.scene_range(total_scenes, world, rank)
        sl = slice(2 * first, 2 * (first + count))
        cfg = R.RenderConfig(n=64, samples=4, shader=tb['shader'], transpose=0, seed=7, scene_begin=2 * first, geom_grad_only=1)
        obj_type, mat, light, cam = t(tb['obj_type']), t(tb['material']), t(tb['light']), t(tb['camera'][sl])
        w2o = t(tb['w2o'][sl])
        target, _, _ = R.render_forward(cfg, obj_type, t(tt['w2o'][sl]), mat, light, cam, None, want_hit=False)
        fn = lambda: R.render_fused_mse(cfg, obj_type, w2o, mat, light, cam, target)
        steps = 50
        ms, _ = timer.run(lambda: [fn() for _ in range(steps)], 3, warm=1, flush=False)
        us = ms / 3 / steps * 1e3
        out_ = dict(us_per_batch=round(us, 1), Mrays_s=round(2.0 * total_scenes * 64 * 64 * 4 / us, 1), scenes_per_gpu=count)
        # the same call replayed from a CUDA graph: a 32-scene shard is a ~27 us launch, less than the ~40 us of
        # Python it takes to enqueue it, so launched from Python the shard is host-bound
        try:
            g, s_ = torch.cuda.CUDAGraph(), torch.cuda.Stream()
            s_.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s_):
                fn()
            torch.cuda.current_stream().wait_stream(s_)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s_):
                fn()
            gms, _ = timer.run(lambda: [g.replay() for _ in range(steps)], 3, warm=1, flush=False)
            gus = gms / 3 / steps * 1e3
            out_.update(us_per_batch_graph_replay=round(gus, 1), Mrays_s_graph_replay=round(2.0 * total_scenes * 64 * 64 * 4 / gus, 1))
        except Exception as e:      # noqa: BLE001
            out_['graph_replay_error'] = repr(e)
            torch.cuda.synchronize()
        return out_
    out['decoder_batch_256_scenes_total'] = decoder(scenes)
    out['decoder_batch_256_scenes_per_gpu'] = decoder(scenes * world)
    spec = importlib.util.spec_from_file_location('orbit_autoencoder', os.path.join(ROOT, 'examples', 'orbit_autoencoder.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for key, total in (('training_step_256_scenes_total', scenes), ('training_step_256_scenes_per_gpu', scenes * world)):
        step, info = mod.make_trainer(total, dev=dev, world=world, rank=rank)
        ms, _ = timer.run(step, 20, warm=3, flush=False)
        info.update(ms_per_step=round(ms / 20, 4), Mrays_s=round(2.0 * total * 64 * 64 * 4 / (ms / 20 * 1e-3) / 1e6, 1),
                    collective='1 NCCL allreduce of the flat encoder-weight gradient + loss per step' if world > 1 else 'none')
        # the same step captured into ONE CUDA graph (encoder, render launch, NCCL allreduce, SGD update)
        gstep, ginfo = mod.make_trainer(total, dev=dev, world=world, rank=rank, graph=True)
        gms, _ = timer.run(gstep, 50, warm=3, flush=False)
        info.update(cuda_graph=dict(captured=ginfo['cuda_graph'], ms_per_step=round(gms / 50, 4),
                                    Mrays_s=round(2.0 * total * 64 * 64 * 4 / (gms / 50 * 1e-3) / 1e6, 1)))
        out[key] = info
    return out


# ------------------------------------------------------------------ GPU leg
def run_b200(args):
    import torch
    import torch.distributed as dist
    from dataclasses import replace
    from reversible_raytracer_b200 import render as R, workloads as W, _native as nat

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: there is no CPU fallback for the product path')
    numa_cpus = bind_to_gpu_numa_node(local) if world > 1 else None
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    nat.lib()  # fail loudly if the extension is not built

    n, S, N = args.n, args.samples, args.objects
    rows_per = (n + world - 1) // world
    rb = min(rank * rows_per, n - 1)
    rc = max(1, min(rows_per, n - rb))
    cfg = R.RenderConfig(n=n, samples=S, shader=nat.SHADER_PHONG, transpose=1, seed=4321, row_begin=rb, row_count=rc,
                         canonical_sweep=int(args.canonical))

    tb = W.stress_tables(N, general=args.general)
    tt = W.stress_tables(N, general=args.general, centre_noise=0.05)
    if args.miss:
        for t_ in (tb, tt):
            t_['w2o'].reshape(-1, 3, 4)[:, 0, 3] -= 1.0e4   # b_x = -c_x/r: shifts every centre far off-axis
    host = {k: torch.from_numpy(tb[k]).pin_memory() for k in ('w2o', 'material', 'light', 'camera')}
    d = {k: v.to(dev) for k, v in host.items()}
    obj_type = torch.from_numpy(tb['obj_type']).to(dev)
    w2o_target = torch.from_numpy(tt['w2o']).to(dev)
    # target slab: the same scene rendered with perturbed centres (resident, like the
    # reference's compiled-in constant `flipped`, match_mirror.py:45)
    target, hit, _ = R.render_forward(cfg, obj_type, w2o_target, d['material'], d['light'], d['camera'], None, want_hit=False)
    _, hit, _ = R.render_forward(cfg, obj_type, d['w2o'], d['material'], d['light'], d['camera'], None, want_hit=True)
    slab_rows = None
    if world > 1 and not args.uniform_slabs:
        # Load balance: the sweep costs the same for every ray, shading + the reverse pass follow the winning rays
        # (138.3 vs 128.2 us per wave with / without hits at 57 % winning rays => a winning ray costs
        # ~0.14 of a ray's sweep on top).  Per-row hit counts of this first render (each rank its uniform slab, one
        # allreduce) -> contiguous slabs of equal estimated cost; every rank computes the same partition.
        from reversible_raytracer_b200 import sharding as Sh
        row_hits = torch.zeros(n, dtype=torch.float64, device=dev)
        row_hits[rb:rb + rc] = (hit >= 0).sum(dim=(0, 2)).to(torch.float64)
        dist.all_reduce(row_hits)
        cost = float(n * S) + 0.14 * row_hits.cpu().numpy()
        def take(slabs_):
            nonlocal rb, rc, cfg, target, hit
            if slabs_[rank] != (rb, rc):
                rb, rc = slabs_[rank]
                cfg = replace(cfg, row_begin=rb, row_count=rc)
                target, _, _ = R.render_forward(cfg, obj_type, w2o_target, d['material'], d['light'], d['camera'], None, want_hit=False)
                _, hit, _ = R.render_forward(cfg, obj_type, d['w2o'], d['material'], d['light'], d['camera'], None, want_hit=True)
        slabs = Sh.balanced_row_slabs(cost, world, align=4)
        take(slabs)
        # one measured correction (still set-up, before warm-up and timing): three fused launches per rank, the
        # per-rank kernel times rescale the cost model of each rank's rows (minus the per-launch constant), and the
        # partition is taken again -- what an optimisation loop would do every few hundred steps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        R.render_fused_mse(cfg, obj_type, d['w2o'], d['material'], d['light'], d['camera'], target, want_image=True)
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        for _ in range(3):
            R.render_fused_mse(cfg, obj_type, d['w2o'], d['material'], d['light'], d['camera'], target, want_image=True)
        e1.record()
        torch.cuda.synchronize()
        t_me = torch.tensor([e0.elapsed_time(e1) / 3.0], dtype=torch.float64, device=dev)
        t_all = [torch.zeros_like(t_me) for _ in range(world)]
        dist.all_gather(t_all, t_me)
        t_all = np.array([float(x) for x in t_all])
        c0 = 0.085                                             # ms per launch that does not scale with the slab
        corr = np.clip((t_all - c0) / max(float(np.mean(t_all)) - c0, 1e-6), 0.8, 1.25)
        for r_, (b_, c_) in enumerate(slabs):
            cost[b_:b_ + c_] *= corr[r_]
        slabs = Sh.balanced_row_slabs(cost, world, align=4)
        take(slabs)
        slab_rows = [c for _, c in slabs]
    hit_rays = torch.tensor([int((hit >= 0).sum())], dtype=torch.float64, device=dev)
    del hit
    timer = Timer(dev, world)
    G = nat.grad_size(N)
    red = torch.zeros(G + 2, dtype=torch.float64, device=dev)
    # the one exchange step: our kernel over NVLink peer memory (sharding.PeerSum ->
    # rrt_peer_allreduce); NCCL allreduce if the symmetric allocation is refused on this box
    peer, collective = None, 'none'
    if world > 1:
        from reversible_raytracer_b200 import sharding as Sh
        ok = torch.ones(1, device=dev)
        try:
            if os.environ.get('RRT_BENCH_NCCL'):
                raise RuntimeError('RRT_BENCH_NCCL set')
            peer = Sh.PeerSum(G, 1, dev)
        except Exception as e:          # noqa: BLE001
            ok.zero_()
            sys.stderr.write('rank %d: PeerSum unavailable (%r), using NCCL\n' % (rank, e))
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok) < 1:
            peer = None
        collective = ('rrt_peer_allreduce: one own kernel per rank over NVLink peer memory, %d float64 per step' % (G + 1)
                      if peer is not None else '1 NCCL allreduce of %d float64 per step' % (G + 2))
        dist.all_reduce(hit_rays)

    def exchange(loss, grad):
        if peer is not None:
            l, g = peer(grad, loss.reshape(1))
            return l[0], g
        red[:G] = grad
        red[G] = loss
        dist.all_reduce(red)            # one NCCL allreduce: gradient vector + loss
        return red[G], red[:G]

    kev = []

    def make_step(c):
        def step():
            if world > 1:
                k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                k0.record()
            loss, grad, _, _ = R.render_fused_mse(c, obj_type, d['w2o'], d['material'], d['light'], d['camera'], target,
                                                   want_image=True)
            if world > 1:
                k1.record()
                kev.append((k0, k1))
                return exchange(loss, grad)
            return loss, grad
        return step
    step = make_step(cfg)

    # The step as an optimisation loop runs it (GDOptimizer captures its step the same way): its launches --
    # rrt_build_records, the memset, the fused render kernel, the peer exchange -- recorded once into ONE CUDA
    # graph and replayed, so that no step waits for Python (after the barrier that opens the timed region the
    # first step would otherwise start with ~100 us of host enqueue time, 5 % of a 2 ms step at 8 GPUs).
    # Falls back to launching from Python if the capture fails.
    graphed = None
    timed_step = step
    if not args.no_graph_step:
        try:
            for _ in range(2):
                step()                                            # eager warm-up (per-stream scratch exists afterwards)
            torch.cuda.synchronize()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            keep_kev = kev
            kev = []                                              # (timing events cannot be recorded inside a capture)

            def captured_body():
                loss, grad, _, _ = R.render_fused_mse(cfg, obj_type, d['w2o'], d['material'], d['light'], d['camera'], target,
                                                       want_image=True)
                return exchange(loss, grad) if world > 1 else (loss, grad)
            with torch.cuda.stream(side):
                captured_body()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_, stream=side):
                g_out = captured_body()
            graphed = g_
            kev = keep_kev

            def timed_step():
                graphed.replay()
                return g_out
        except Exception as e:          # noqa: BLE001
            sys.stderr.write('rank %d: CUDA-graph capture of the step failed (%r); launching from Python\n' % (rank, e))
            torch.cuda.synchronize()
            graphed, timed_step = None, step

    # everything with host-side start-up cost (NVML init, event creation) happens BEFORE the
    # barrier, so that all ranks enter the timed region together
    sampler = ClockSampler(local)
    sampler.start()
    total_ms, per_step = timer.run(timed_step, args.steps, warm=max(args.warmup, 3))
    sampler.stop_flag = True
    kernel_ms = None
    if world > 1:
        if graphed is not None:     # per-rank render time (events cannot be recorded inside the graph): the same K steps
            kev = []                # launched from Python, same flush / barrier protocol, for the record only
            timer.run(step, args.steps, warm=1)
        km = torch.tensor([sum(a.elapsed_time(b) for a, b in kev[-args.steps:]) / args.steps], dtype=torch.float64, device=dev)
        gathered = [torch.zeros_like(km) for _ in range(world)]
        dist.all_gather(gathered, km)
        kernel_ms = [round(float(x), 3) for x in gathered]
    rays = float(n) * n * S
    value = rays * args.steps / (total_ms * 1e-3) / 1e6
    ms_step = total_ms / args.steps

    # ---- the same step with the other sweep (canonical <-> pre-filter), for the roofline accounting
    other = replace(cfg, canonical_sweep=1 - cfg.canonical_sweep)
    nother = min(args.steps, 4)
    other_ms, _ = timer.run(make_step(other), nother, warm=1)
    other_ms /= nother

    # ---- forward only (BASELINE.json asks for forward AND forward+backward at every N): the same
    # slab rendered by rrt_render_forward, no exchange needed; max over ranks
    def fwd_step():
        return R.render_forward(cfg, obj_type, d['w2o'], d['material'], d['light'], d['camera'], None, want_hit=False)
    nfwd = min(args.steps, 5)
    fwd_ms, _ = timer.run(fwd_step, nfwd, warm=1)
    fwd_ms /= nfwd

    # ---- end-to-end through the public functional API with HOST buffers: every step
    # uploads the scene-parameter tables from pinned memory and reads back loss + gradient
    pin_grad = torch.empty(G, dtype=torch.float32).pin_memory()
    pin_loss = torch.empty(1, dtype=torch.float64).pin_memory()

    def finish(loss, grad):
        if world > 1:
            l, g = exchange(loss, grad)
            pin_grad.copy_(g.float(), non_blocking=True)
            pin_loss.copy_(l.reshape(1), non_blocking=True)
        else:
            pin_grad.copy_(grad, non_blocking=True)
            pin_loss.copy_(loss.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(pin_loss[0])

    def e2e_step():
        dd = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        loss, grad, _, _ = R.render_fused_mse(cfg, obj_type, dd['w2o'], dd['material'], dd['light'], dd['camera'], target,
                                               want_image=True)
        return finish(loss, grad)

    e_ms, _ = timer.run(e2e_step, args.steps, warm=1)
    e2e_value = rays * args.steps / (e_ms * 1e-3) / 1e6
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = G * 4 + 8

    # The full host-buffer variant: the target slab ALSO comes from pinned host memory every
    # step and the rendered image slab is read back (2 x 12 B/pixel over PCIe).  Measured twice:
    # plainly (copy, kernel, copy back to back on one stream) and through
    # render.StreamedFusedMSE (row slabs pipelined over copy-in / kernel / copy-out streams).
    pin_target = target.cpu().pin_memory()
    pin_image = torch.empty_like(pin_target).pin_memory()
    streamed = R.StreamedFusedMSE(cfg, N, dev, want_image=True)

    def e2e_full_step():
        tgt = pin_target.to(dev, non_blocking=True)
        dd = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        loss, grad, img, _ = R.render_fused_mse(cfg, obj_type, dd['w2o'], dd['material'], dd['light'], dd['camera'], tgt,
                                                 want_image=True)
        pin_image.copy_(img, non_blocking=True)
        return finish(loss, grad)

    def e2e_streamed_step():
        dd = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        loss, grad = streamed(obj_type, dd['w2o'], dd['material'], dd['light'], dd['camera'], pin_target, pin_image)
        return finish(loss, grad)

    nfull = min(args.steps, 5)
    f_ms, _ = timer.run(e2e_full_step, nfull, warm=1)
    e2e_full_value = rays * nfull / (f_ms * 1e-3) / 1e6
    s_ms, _ = timer.run(e2e_streamed_step, args.steps, warm=1)
    e2e_streamed_value = rays * args.steps / (s_ms * 1e-3) / 1e6
    full_bytes = pin_target.numel() * 4

    # ---- once, outside every timed region: the streamed host-buffer path must reproduce the resident
    # launch -- image bit for bit, loss and gradient up to the summation order of the slabs
    loss_r, grad_r, img_r, _ = R.render_fused_mse(cfg, obj_type, d['w2o'], d['material'], d['light'], d['camera'], target,
                                                   want_image=True)
    pin_image.zero_()
    loss_s, grad_s = streamed(obj_type, d['w2o'], d['material'], d['light'], d['camera'], pin_target, pin_image)
    torch.cuda.synchronize()
    e2e_check = dict(image_bits_identical=bool(torch.equal(pin_image, img_r.cpu())),
                     loss_rel_diff=abs(float(loss_s) - float(loss_r)) / max(abs(float(loss_r)), 1e-30),
                     grad_max_rel_diff=float((grad_s - grad_r).abs().max() / grad_r.abs().max()))
    flags = torch.tensor([1.0 if e2e_check['image_bits_identical'] else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    e2e_check['image_bits_identical_all_ranks'] = bool(float(flags) > 0)
    assert e2e_check['image_bits_identical_all_ranks'], 'streamed e2e path does not reproduce the resident image'
    assert e2e_check['loss_rel_diff'] < 1e-6 and e2e_check['grad_max_rel_diff'] < 1e-4, e2e_check

    # ---- driver-visible proof of the exchange (N > 1): the same [grad, loss] through our peer-memory
    # kernel and through an NCCL allreduce; identical bits on every rank; and against ONE GPU
    # rendering the whole image (the workload is deterministic up to float summation order)
    exchange_check = None
    if world > 1:
        loss_l, grad_l, _, _ = R.render_fused_mse(cfg, obj_type, d['w2o'], d['material'], d['light'], d['camera'], target)
        l_x, g_x = exchange(loss_l, grad_l)
        mine = torch.cat([g_x.double().reshape(-1), l_x.double().reshape(1)]).clone()
        ref = torch.cat([grad_l.double().reshape(-1), loss_l.double().reshape(1)])
        dist.all_reduce(ref)                                    # NCCL on the same inputs
        root = mine.clone()
        dist.broadcast(root, 0)
        same = torch.tensor([1.0 if torch.equal(root, mine) else 0.0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        exchange_check = dict(max_abs_diff_vs_nccl=float((mine - ref).abs().max()),
                              max_rel_diff_vs_nccl=float((mine - ref).abs().max() / ref.abs().max()),
                              bits_identical_across_ranks=bool(float(same) > 0),
                              via='rrt_peer_allreduce' if peer is not None else 'nccl')
        if rank == 0:                                           # the unsharded image on one GPU
            cfg1 = replace(cfg, row_begin=0, row_count=0)
            tgt1, _, _ = R.render_forward(cfg1, obj_type, w2o_target, d['material'], d['light'], d['camera'], None, want_hit=False)
            l1, g1, _, _ = R.render_fused_mse(cfg1, obj_type, d['w2o'], d['material'], d['light'], d['camera'], tgt1)
            exchange_check['loss_rel_diff_vs_one_gpu'] = abs(float(l1) - float(l_x)) / abs(float(l1))
            exchange_check['grad_max_rel_diff_vs_one_gpu'] = float((g1.double() - g_x.double()).abs().max() / g1.abs().max())
            del tgt1
        dist.barrier()

    # ---- roofline (every N): FP32 FMA pipe.  Rank 0 measures the denominator on its own GPU.
    out = None
    c4 = None
    if world > 1 and not args.no_extras and not args.general:
        try:
            c4 = c4_sharded(dev, world, rank, timer)
        except Exception as e:      # noqa: BLE001
            c4 = {'error': repr(e)}
    cull_ms = None
    if world > 1 and not args.no_extras:
        ccfg = replace(cfg, cull=1)
        cm, _ = timer.run(make_step(ccfg), 5, warm=1)
        cull_ms = cm / 5
    if rank == 0:
        flops = W.algorithmic_flops(rays, N, float(hit_rays), general=args.general)
        peak_tf, _ = R.measure_fp32_peak(1, 4096)
        peak_scalar, _ = R.measure_fp32_peak(0, 4096)
        peak_mixed, _ = R.measure_fp32_peak(2, 4096)
        peak_all = peak_tf * world                                # N identical GPUs
        ach = flops / (ms_step * 1e-3) / 1e12
        prefilter_ms, canonical_ms = (other_ms, ms_step) if args.canonical else (ms_step, other_ms)
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, 'profiles', 'traffic_fused_c5.json')))
            if (n, S, N, args.general, world) == (4096, 4, 1024, False, 1):
                traffic = tj['dram_bytes_read'] + tj['dram_bytes_write']
                traffic_src = tj['source']
        except Exception:
            pass
        # executed FMA-lane work of the default sweep: 5 fused multiply-adds per (ray, sphere) pair
        # (10 flops) instead of the 11 operations / 16 credited flops of the reference arithmetic
        exec_flops = rays * N * 10.0 + float(hit_rays) * 320.0
        roof = dict(bound='fp32', achieved=ach, peak=peak_all, unit='TFLOP/s', frac=ach / peak_all, traffic=traffic,
                    traffic_unit='bytes per launch (dram read+write)', traffic_source=traffic_src,
                    kernel='render_kernel<2,4,FUSED>', algorithmic_flops=flops, n_gpus=world,
                    peak_per_gpu=peak_tf,
                    sweep='canonical (RRT_FLAG_CANONICAL_SWEEP)' if args.canonical else
                          'conservative pre-filter + canonical decision (default)',
                    note='`achieved` = CREDITED algorithmic flops (SURVEY.md 8d: 16 per diagonal ray-sphere test, one sweep, '
                         '+ 320 per winning ray) / step time.  The default sweep decides 99.9 % of the pairs with a 5-FMA '
                         'float32 quadratic form that conservatively bounds the sign of the reference discriminant and '
                         'hands the rest to the canonical arithmetic (bit-identical masks, tested), so `frac` may exceed '
                         'the 0.727 ceiling of the canonical arithmetic (16 credited flops per 22 FMA-lane-ops) and even 1.0 '
                         '(credited flops are those of the reference arithmetic; 10 per pair are executed); '
                         '`executed` is the FMA-pipe work actually issued, `canonical_sweep` the same step with every '
                         'pair in the reference arithmetic.',
                    executed=dict(flops=exec_flops, tflops=exec_flops / (prefilter_ms * 1e-3) / 1e12,
                                  frac=exec_flops / (prefilter_ms * 1e-3) / 1e12 / peak_all,
                                  note='10 executed flops per pair: 5 FFMA2-lane FMAs; ceiling of this loop = 160 FFMA2 '
                                       'issue cycles of 190 per 4 objects = 0.84'),
                    canonical_sweep=dict(ms_per_step=canonical_ms, Mrays_s=rays / (canonical_ms * 1e-3) / 1e6,
                                         achieved=flops / (canonical_ms * 1e-3) / 1e12,
                                         frac=flops / (canonical_ms * 1e-3) / 1e12 / peak_all, ceiling=16.0 / 22.0),
                    prefilter_sweep=dict(ms_per_step=prefilter_ms, Mrays_s=rays / (prefilter_ms * 1e-3) / 1e6,
                                         achieved=flops / (prefilter_ms * 1e-3) / 1e12,
                                         frac=flops / (prefilter_ms * 1e-3) / 1e12 / peak_all),
                    issue_model=dict(
                        ffma2_only_tflops=peak_tf, scalar_ffma_tflops=peak_scalar, ffma2_plus_1_alu_per_4_tflops=peak_mixed,
                        note='an FFMA2 holds the SMSP issue port for 2 cycles (nothing issues in its shadow), so every '
                             'non-FMA instruction of the sweep costs FP32 throughput'),
                    peak_source='measured in this run on rank 0 (x n_gpus): packed FFMA2 micro-benchmark '
                                '(librrt_b200_bench.so, SM count from the device); MEASURED_PEAKS.json has no FP32 entry; '
                                'theoretical 148*128*2*1.965 GHz = 74.45 per GPU',
                    hbm=dict(algorithmic_bytes=float(n) * n * 24,
                             achieved_gbs=float(n) * n * 24 / (ms_step * 1e-3) / 1e9))
        hbm_gbs = 6534.8
        try:
            hbm_gbs = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
        except Exception:
            pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(args, args.cpu_seconds)
        out = dict(metric='Mrays/s fwd+bwd', value=value, unit='Mrays/s', n_gpus=world, steps=args.steps,
                   warmup=max(args.warmup, 3), ms_per_step=ms_step, higher_is_better=True, scaling='strong',
                   vs_baseline=None, dtype='f32', data='synthetic',
                   config=dict(workload=WORKLOAD if not args.general else WORKLOAD.replace('translate*scale', 'translate*rotate*scale'),
                               n=n, samples=S, objects=N,
                               sharding=('row slabs, %d rows per GPU' % rows_per) if slab_rows is None else
                               'contiguous row slabs of equal estimated cost (rays + 0.14 x winning rays per row from a first '
                               'render, corrected once by measured per-rank kernel times; set-up, before warm-up): rows per GPU %s' % slab_rows,
                               collective=collective,
                               l2='256 MiB flush write between timed iterations (outside the timed intervals)',
                               host_cpus_bound_to_gpu_numa_node=numa_cpus,
                               jitter='in-kernel counter RNG, seed 4321',
                               step_launch='one CUDA-graph replay per step' if graphed is not None else 'launched from Python'),
                   e2e=dict(value=e2e_streamed_value, unit='Mrays/s', h2d_bytes_per_step=h2d + full_bytes,
                            d2h_bytes_per_step=d2h + full_bytes,
                            note='every step: scene-parameter tables AND the target slab come from pinned host memory, '
                                 'loss + gradient vector AND the rendered image slab go back to pinned host memory; '
                                 'render.StreamedFusedMSE pipelines %d row slabs over copy-in / kernel / copy-out streams' % len(streamed.bounds),
                            check_vs_resident_launch=e2e_check,
                            launches_per_step=streamed.launches_per_call + (1 if peer is not None else 0),
                            same_buffers_unpipelined=dict(value=e2e_full_value, unit='Mrays/s'),
                            parameters_only=dict(
                                value=e2e_value, unit='Mrays/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                                note='target stays resident like the reference\'s compiled-in constant '
                                     '(match_mirror.py:45); only the parameter tables go up and loss + gradients come back')),
                   # per step: rrt_build_records (>= 64 objects), the fused render kernel (gradient finalisation
                   # folded in through rrt_scene.ticket), and at N > 1 the peer-memory exchange kernel
                   gpu_launches=(1 + (1 if N >= R.RECORDS_MIN_N else 0) + (1 if peer is not None else 0)) * args.steps,
                   clocks=sampler.summary())
        out['forward_only'] = dict(ms_per_step=fwd_ms, value=float(n) * n * S / (fwd_ms * 1e-3) / 1e6, unit='Mrays/s')
        # forward credits the sweep + 64 nominal flops per winning ray
        f_flops = rays * N * (28.0 if args.general else 16.0) + float(hit_rays) * 64.0
        out['forward_only']['frac_fp32_peak'] = f_flops / (fwd_ms * 1e-3) / 1e12 / peak_all
        out['roofline'] = roof
        if exchange_check is not None:
            out['exchange_check'] = exchange_check
        if kernel_ms is not None:
            out['rank0_step_ms'] = [round(x, 3) for x in per_step]
            out['per_rank_render_ms'] = kernel_ms     # fused kernel per rank, before the exchange
        if world == 1 and not args.no_extras:
            try:
                out['other_configs'] = other_configs(dev, peak_tf, hbm_gbs, with_cpu=not args.no_cpu_baseline)
            except Exception as e:          # extras must never take the headline down
                out['other_configs'] = {'error': repr(e)}
        if world > 1 and not args.no_extras:
            oc_ = {}
            if c4 is not None:
                oc_['C4_orbit_autoencoder_sharded'] = c4
            if cull_ms is not None:
                oc_['C5_fused_with_culling'] = dict(
                    ms=round(cull_ms, 3), Mrays_s=round(rays / (cull_ms * 1e-3) / 1e6, 1),
                    note='RRT_FLAG_CULL on the same row slabs: work follows where the spheres are, so slabs are no '
                         'longer perfectly balanced; same bits as the exhaustive sweep; never a roofline fraction')
            out['other_configs'] = oc_
        if cpu is not None:
            out['cpu_baseline'] = cpu
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_c4(args):
    """BASELINE config 4 (test_optimization.py:17-44, autoencoder_2ly.py:82-91 scaled to a batch):
    `--scenes` orbit scenes x 2 camera views, 64x64, S=4, fused forward + squared error + reverse
    pass.  Scene ranges are sharded across ranks (sharding.scene_range); per-scene gradients stay
    on their rank (they feed the local encoder backward), so the render path has no collective."""
    import torch
    import torch.distributed as dist
    from reversible_raytracer_b200 import _native as nat
    world, rank = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    nat.lib()
    timer = Timer(dev, world)
    res = c4_sharded(dev, world, rank, timer, args.scenes)
    if rank == 0:
        dec = res['decoder_batch_256_scenes_total']
        print(json.dumps(dict(
            metric='Mrays/s fwd+bwd', value=dec['Mrays_s'], unit='Mrays/s', n_gpus=world,
            steps=150, warmup=50, ms_per_step=dec['us_per_batch'] / 1e3, higher_is_better=True,
            scaling='strong', vs_baseline=None, dtype='f32', data='synthetic',
            config=dict(workload='C4 orbit autoencoder decoder batch: %d scenes x 2 views, 64x64, S=4, 2 spheres, '
                                 'fused fwd+mse+bwd (d/d w2o only)' % args.scenes, sharding='%d scenes per GPU' % dec['scenes_per_gpu'],
                        collective='none on the render path (per-scene gradients stay local)',
                        l2='working set is L2-resident by nature of the workload'),
            gpu_launches=150, details=res)))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    elif args.workload == 'c4':
        run_c4(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
