#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native differentiable ray tracer.

Metric (BASELINE.json): Mrays/s forward+backward on the synthetic stress scene C5
(4096 x 4096 image, 4 anti-alias samples, 1024 spheres), one "step" = one fused
forward + squared-error loss + reverse pass over the whole image; with N GPUs the
image's row slabs are sharded across ranks (total work fixed => strong scaling)
and the small parameter-gradient vector + loss are summed by ONE kernel of our own over
NVLink peer memory (rrt_peer_allreduce; NCCL allreduce as the fallback).

    python bench.py [--gpus N] [--steps K] [--warmup W]            this framework
    python bench.py --impl reference [...]                         CPU restatement of the
        reference algorithm (oracle/oracle_c.c, all host threads) on a bounded sample
        of the same workload -- the reference itself is Python 2 + Theano and cannot
        run in this image.

Prints ONE JSON line on rank 0.
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
    ap.add_argument('--cpu-seconds', type=float, default=20.0)
    ap.add_argument('--workload', default='c5', choices=['c5', 'c4'],
                    help="c5 (default): the headline stress scene; c4: BASELINE config 4, the orbit autoencoder's "
                         "decoder batch (256 scenes x 2 views, 64x64, S=4), scene ranges sharded across ranks")
    ap.add_argument('--scenes', type=int, default=256)
    return ap.parse_args()


# ------------------------------------------------------------------ CPU restatement leg
def cpu_fused_sample(args, rows, row_begin=None):
    """Times oracle_c's fused fwd+mse+bwd on `rows` rows of the workload (all host
    threads).  The ONLY place bench.py executes anything under oracle/."""
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


# ------------------------------------------------------------------ the other BASELINE configs
def other_configs(dev):
    """Optimise-step latency of the reference's own small configs (C1 optimize_brightness.py,
    C3 match_mirror.py: 128x128, S=4) through the drop-in API + GDOptimizer (CUDA-graph
    replayed step incl. the device->host read of the loss), throughput of the batched
    autoencoder decoder workload C4 (256 scenes x 2 views, 64x64, S=4, fused fwd+mse+bwd),
    and C5g (general affine spheres) / C5 forward-only for the record."""
    import torch
    from reversible_raytracer_b200 import render as R, workloads as W, _native as nat
    from tools import latency as L
    out = {}
    train, _ = L.c1()
    out['C1_optimize_brightness_step_us'] = round(L.timeit(train, warm=6, iters=100), 1)
    train, _ = L.c3(False)
    out['C3_match_mirror_step_us'] = round(L.timeit(train, warm=6, iters=100), 1)
    train, _ = L.c3('graph')
    out['C3_match_mirror_fused_step_us'] = round(L.timeit(train, warm=6, iters=100), 1)
    train, _ = L.c3(True)                       # Scene.mse_cost: whole step = one kernel launch (rrt_small_step_mse)
    if train.state.get('whole_step') is not None:
        out['C3_match_mirror_whole_step_kernel_us'] = round(L.timeit(train, warm=6, iters=100), 1)
    fn, rays = L.c4(256)
    us = L.timeit(fn, warm=5, iters=100)
    out['C4_orbit_256x2_fused'] = dict(us_per_batch=round(us, 1), Mrays_s=round(rays / us, 1))

    def stress(general, fwd_only, samples=4, n=4096, N=1024, iters=3, cull=0):
        tb = W.stress_tables(N, general=general)
        t = lambda a: torch.from_numpy(a).to(dev)
        cfg = R.RenderConfig(n=n, samples=samples, shader=nat.SHADER_PHONG, transpose=1, seed=4321, cull=cull)
        args = (t(tb['obj_type']), t(tb['w2o']), t(tb['material']), t(tb['light']), t(tb['camera']))
        target, _, _ = R.render_forward(cfg, *args, None, want_hit=False)
        fn = (lambda: R.render_forward(cfg, *args, None, want_hit=False)) if fwd_only else \
             (lambda: R.render_fused_mse(cfg, *args, target, want_image=True))
        us = L.timeit(fn, warm=2, iters=iters)
        return dict(ms=round(us / 1e3, 3), Mrays_s=round(n * n * samples / us, 1))
    out['C5_forward_only'] = stress(False, True)
    out['C5_S1_fused'] = stress(False, False, samples=1)
    g = stress(True, False)
    g['frac_fp32_peak_28flop_per_test'] = None
    out['C5g_general_affine_fused'] = g
    # NOT roofline-accountable: conservative per-tile culling skips work (results bit-identical, tested)
    c = stress(False, False, cull=1, iters=5)
    c['note'] = 'RRT_FLAG_CULL: same bits as the exhaustive sweep, reported separately, never as a roofline fraction'
    out['C5_fused_with_culling'] = c
    return out


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


# ------------------------------------------------------------------ GPU leg
def run_b200(args):
    import torch
    import torch.distributed as dist
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
    cfg = R.RenderConfig(n=n, samples=S, shader=nat.SHADER_PHONG, transpose=1, seed=4321, row_begin=rb, row_count=rc)

    tb = W.stress_tables(N, general=args.general)
    tt = W.stress_tables(N, general=args.general, centre_noise=0.05)
    if args.miss:
        for t_ in (tb, tt):
            t_['w2o'].reshape(-1, 3, 4)[:, 0, 3] -= 1.0e4   # b_x = -c_x/r: shifts every centre far off-axis
    host = {k: torch.from_numpy(tb[k]).pin_memory() for k in ('w2o', 'material', 'light', 'camera')}
    d = {k: v.to(dev) for k, v in host.items()}
    obj_type = torch.from_numpy(tb['obj_type']).to(dev)
    # target slab: the same scene rendered with perturbed centres (resident, like the
    # reference's compiled-in constant `flipped`, match_mirror.py:45)
    target, hit, _ = R.render_forward(cfg, obj_type, torch.from_numpy(tt['w2o']).to(dev), d['material'], d['light'],
                                      d['camera'], None, want_hit=False)
    _, hit, _ = R.render_forward(cfg, obj_type, d['w2o'], d['material'], d['light'], d['camera'], None, want_hit=True)
    hit_rays = torch.tensor([int((hit >= 0).sum())], dtype=torch.float64, device=dev)
    del hit
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
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

    def exchange(loss, grad):
        if peer is not None:
            l, g = peer(grad, loss.reshape(1))
            return l[0], g
        red[:G] = grad
        red[G] = loss
        dist.all_reduce(red)            # one NCCL allreduce: gradient vector + loss
        return red[G], red[:G]

    kev = []

    def step():
        if world > 1:
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
        loss, grad, _, _ = R.render_fused_mse(cfg, obj_type, d['w2o'], d['material'], d['light'], d['camera'], target,
                                               want_image=True)
        if world > 1:
            k1.record()
            kev.append((k0, k1))
            return exchange(loss, grad)
        return loss, grad

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    # everything with host-side start-up cost (NVML init, event creation) happens BEFORE the
    # barrier, so that all ranks enter the timed region together
    sampler = ClockSampler(local)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world > 1:
        dist.all_reduce(hit_rays)
        torch.cuda.synchronize()
        dist.barrier()
    torch.cuda.synchronize()
    for k in range(args.steps):
        flush.fill_(k & 0xff)               # L2 flush between timed iterations (outside the timed intervals)
        evs[k][0].record()
        step()
        evs[k][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    kernel_ms = None
    if world > 1:
        km = torch.tensor([sum(a.elapsed_time(b) for a, b in kev[-args.steps:]) / args.steps], dtype=torch.float64, device=dev)
        gathered = [torch.zeros_like(km) for _ in range(world)]
        dist.all_gather(gathered, km)
        kernel_ms = [round(float(x), 3) for x in gathered]
    tms = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    total_ms = float(tms)
    rays = float(n) * n * S
    value = rays * args.steps / (total_ms * 1e-3) / 1e6

    # ---- forward only (BASELINE.json asks for forward AND forward+backward at every N): the same
    # slab rendered by rrt_render_forward, no exchange needed; max over ranks
    def fwd_step():
        return R.render_forward(cfg, obj_type, d['w2o'], d['material'], d['light'], d['camera'], None, want_hit=False)
    fwd_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    f_ms_total, nfwd = 0.0, min(args.steps, 5)
    for k in range(nfwd):
        flush.fill_(k & 0xff)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fwd_step()
        b.record()
        torch.cuda.synchronize()
        f_ms_total += a.elapsed_time(b)
    fwd_t = torch.tensor([f_ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(fwd_t, op=dist.ReduceOp.MAX)
    fwd_ms = float(fwd_t) / nfwd

    # ---- end-to-end through the public functional API with HOST buffers: every step
    # uploads the scene-parameter tables from pinned memory and reads back loss + gradient
    pin_grad = torch.empty(G, dtype=torch.float32).pin_memory()
    pin_loss = torch.empty(1, dtype=torch.float64).pin_memory()

    def e2e_step():
        dd = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        loss, grad, _, _ = R.render_fused_mse(cfg, obj_type, dd['w2o'], dd['material'], dd['light'], dd['camera'], target,
                                               want_image=True)
        if world > 1:
            l, g = exchange(loss, grad)
            pin_grad.copy_(g.float(), non_blocking=True)
            pin_loss.copy_(l.reshape(1), non_blocking=True)
        else:
            pin_grad.copy_(grad, non_blocking=True)
            pin_loss.copy_(loss.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(pin_loss[0])

    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e_ms = 0.0
    for k in range(args.steps):
        flush.fill_(k & 0xff)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        e2e_step()
        b.record()
        torch.cuda.synchronize()
        e_ms += a.elapsed_time(b)
    ems = torch.tensor([e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_value = rays * args.steps / (float(ems) * 1e-3) / 1e6
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = G * 4 + 8

    # The full host-buffer variant: the target slab ALSO comes from pinned host memory every
    # step and the rendered image slab is read back (2 x 12 B/pixel over PCIe).  Measured twice:
    # plainly (copy, kernel, copy back to back on one stream) and through
    # render.StreamedFusedMSE (row slabs pipelined over copy-in / kernel / copy-out streams).
    pin_target = target.cpu().pin_memory()
    pin_image = torch.empty_like(pin_target).pin_memory()
    streamed = R.StreamedFusedMSE(cfg, N, dev, want_image=True)

    def finish(loss, grad):
        if world > 1:
            l, g = exchange(loss, grad)
            pin_grad.copy_(g.float(), non_blocking=True)
            pin_loss.copy_(l.reshape(1), non_blocking=True)
        else:
            pin_grad.copy_(grad, non_blocking=True)
            pin_loss.copy_(loss.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def e2e_full_step():
        tgt = pin_target.to(dev, non_blocking=True)
        dd = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        loss, grad, img, _ = R.render_fused_mse(cfg, obj_type, dd['w2o'], dd['material'], dd['light'], dd['camera'], tgt,
                                                 want_image=True)
        pin_image.copy_(img, non_blocking=True)
        finish(loss, grad)

    def e2e_streamed_step():
        dd = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        loss, grad = streamed(obj_type, dd['w2o'], dd['material'], dd['light'], dd['camera'], pin_target, pin_image)
        finish(loss, grad)

    def time_e2e(fn, steps):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = 0.0
        for k in range(steps):
            flush.fill_(k & 0xff)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ms += a.elapsed_time(b)
        t_ = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return rays * steps / (float(t_) * 1e-3) / 1e6

    nfull = min(args.steps, 5)
    e2e_full_value = time_e2e(e2e_full_step, nfull)
    e2e_streamed_value = time_e2e(e2e_streamed_step, args.steps)
    full_bytes = pin_target.numel() * 4

    out = None
    if rank == 0:
        flops = W.algorithmic_flops(rays, N, float(hit_rays), general=args.general)
        ms_step = total_ms / args.steps
        roof = None
        cpu = None
        if world == 1:
            peak_tf, _ = R.measure_fp32_peak(1, 4096)
            peak_scalar, _ = R.measure_fp32_peak(0, 4096)
            peak_mixed, _ = R.measure_fp32_peak(2, 4096)
            ach = flops / (ms_step * 1e-3) / 1e12
            traffic, traffic_src = None, None
            try:
                tj = json.load(open(os.path.join(ROOT, 'profiles', 'traffic_fused_c5.json')))
                if (n, S, N, args.general) == (4096, 4, 1024, False):
                    traffic = tj['dram_bytes_read'] + tj['dram_bytes_write']
                    traffic_src = tj['source']
            except Exception:
                pass
            roof = dict(bound='fp32', achieved=ach, peak=peak_tf, unit='TFLOP/s', frac=ach / peak_tf, traffic=traffic,
                        traffic_unit='bytes per launch (dram read+write)', traffic_source=traffic_src,
                        kernel='render_kernel<2,4,FUSED>', algorithmic_flops=flops,
                        issue_model=dict(
                            ffma2_only_tflops=peak_tf, scalar_ffma_tflops=peak_scalar, ffma2_plus_1_alu_per_4_tflops=peak_mixed,
                            note='an FFMA2 holds the SMSP issue port for 2 cycles (nothing issues in its shadow), so every '
                                 'non-FMA instruction of the sweep costs FP32 throughput; the canonical order itself caps '
                                 'at 16 credited flops / 22 FMA-lane-ops = 72.7% of peak'),
                        peak_source='measured in this run: packed FFMA2 micro-benchmark (rrt_measure_fp32_peak); '
                                    'MEASURED_PEAKS.json has no FP32 entry; theoretical 148*128*2*1.965 GHz = 74.45',
                        hbm=dict(algorithmic_bytes=float(n) * n * 24,
                                 achieved_gbs=float(n) * n * 24 / (ms_step * 1e-3) / 1e9))
            if not args.no_cpu_baseline:
                cpu = cpu_baseline(args, args.cpu_seconds)
        out = dict(metric='Mrays/s fwd+bwd', value=value, unit='Mrays/s', n_gpus=world, steps=args.steps,
                   warmup=max(args.warmup, 3), ms_per_step=ms_step, higher_is_better=True, scaling='strong',
                   vs_baseline=None, dtype='f32', data='synthetic',
                   config=dict(workload=WORKLOAD if not args.general else WORKLOAD.replace('translate*scale', 'translate*rotate*scale'),
                               n=n, samples=S, objects=N, sharding='row slabs, %d rows per GPU' % rows_per,
                               collective=collective,
                               l2='256 MiB flush write between timed iterations (outside the timed intervals)',
                               host_cpus_bound_to_gpu_numa_node=numa_cpus,
                               jitter='in-kernel counter RNG, seed 4321'),
                   e2e=dict(value=e2e_streamed_value, unit='Mrays/s', h2d_bytes_per_step=h2d + full_bytes,
                            d2h_bytes_per_step=d2h + full_bytes,
                            note='every step: scene-parameter tables AND the target slab come from pinned host memory, '
                                 'loss + gradient vector AND the rendered image slab go back to pinned host memory; '
                                 'render.StreamedFusedMSE pipelines %d row slabs over copy-in / kernel / copy-out streams' % len(streamed.bounds),
                            same_buffers_unpipelined=dict(value=e2e_full_value, unit='Mrays/s'),
                            parameters_only=dict(
                                value=e2e_value, unit='Mrays/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                                note='target stays resident like the reference\'s compiled-in constant '
                                     '(match_mirror.py:45); only the parameter tables go up and loss + gradients come back')),
                   # per step: rrt_build_records (>= 64 objects), the fused render kernel, finalize_grads,
                   # and at N > 1 the peer-memory exchange kernel
                   gpu_launches=(2 + (1 if N >= R.RECORDS_MIN_N else 0) + (1 if peer is not None else 0)) * args.steps,
                   clocks=sampler.summary())
        out['forward_only'] = dict(ms_per_step=fwd_ms, value=float(n) * n * S / (fwd_ms * 1e-3) / 1e6, unit='Mrays/s',
                                   frac_fp32_peak=None)
        if kernel_ms is not None:
            out['rank0_step_ms'] = [round(a.elapsed_time(b), 3) for a, b in evs]
            out['per_rank_render_ms'] = kernel_ms     # fused kernel + finalize per rank, before the allreduce
        if roof is not None:
            out['roofline'] = roof
            # forward credits the sweep + 64 nominal flops per winning ray
            f_flops = rays * N * (28.0 if args.general else 16.0) + float(hit_rays) * 64.0
            out['forward_only']['frac_fp32_peak'] = f_flops / (fwd_ms * 1e-3) / 1e12 / roof['peak']
        if world == 1 and not args.no_extras:
            try:
                oc_ = other_configs(dev)
                g_ = oc_.get('C5g_general_affine_fused')
                if g_ and roof is not None:
                    g_['frac_fp32_peak_28flop_per_test'] = round(
                        (rays * N * 28) / (g_['ms'] * 1e-3) / 1e12 / roof['peak'], 4)
                out['other_configs'] = oc_
            except Exception as e:          # extras must never take the headline down
                out['other_configs'] = {'error': repr(e)}
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
    from reversible_raytracer_b200 import render as R, workloads as W, sharding as Sh, _native as nat
    world, rank = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    nat.lib()
    tb, tt = W.orbit_tables(args.scenes), W.orbit_tables(args.scenes, centre_noise=0.5)
    first, count = Sh.scene_range(args.scenes, world, rank)
    sl = slice(2 * first, 2 * (first + count))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cfg = R.RenderConfig(n=64, samples=4, shader=tb['shader'], transpose=0, seed=7, scene_begin=2 * first)
    obj_type, mat, light, cam = t(tb['obj_type']), t(tb['material']), t(tb['light']), t(tb['camera'][sl])
    w2o = t(tb['w2o'][sl])
    target, _, _ = R.render_forward(cfg, obj_type, t(tt['w2o'][sl]), mat, light, cam, None, want_hit=False)
    step = lambda: R.render_fused_mse(cfg, obj_type, w2o, mat, light, cam, target)
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        step()
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    rays = 2.0 * args.scenes * 64 * 64 * 4
    if rank == 0:
        print(json.dumps(dict(
            metric='Mrays/s fwd+bwd', value=rays * args.steps / (float(ms) * 1e-3) / 1e6, unit='Mrays/s', n_gpus=world,
            steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=float(ms) / args.steps, higher_is_better=True,
            scaling='strong', vs_baseline=None, dtype='f32', data='synthetic',
            config=dict(workload='C4 orbit autoencoder decoder batch: %d scenes x 2 views, 64x64, S=4, 2 spheres, '
                                 'fused fwd+mse+bwd' % args.scenes, sharding='%d scenes per GPU' % count,
                        collective='none on the render path (per-scene gradients stay local)',
                        l2='working set (%.0f MB per rank) is L2-resident by nature of the workload' % (2 * count * 64 * 64 * 3 * 4 * 2 / 1e6)),
            gpu_launches=2 * args.steps)))
    if world > 1:
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
