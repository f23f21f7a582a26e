#!/usr/bin/env python
"""Benchmark of the shapegan hot path on B200 (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload wgan|wgan_gp|autodecoder]

One "step" of the default workload = BASELINE.json configs[1]: the 3D-CNN WGAN G+D step on 32^3 voxels, batch 64 per
GPU, bf16 -- exactly train_wgan.py:62-71 (critic update: fake B + real B, RMSprop, clip) + :75-84 (generator update).
metric = voxels/s = n_gpus * B * 32^3 / step time.  Prints ONE JSON line on rank 0.

  value     : step timed with inputs resident in HBM (CUDA events per iteration, L2 flushed between iterations)
  e2e       : same step through the public API with the batch + latents copied from pinned host memory and the two
              loss scalars read back, every step, inside the timed region
  roofline  : the dominant kernel (tcgen05 implicit GEMM, Conv3d 64->128 forward at B=64) timed alone with CUDA events
              against MEASURED_PEAKS.json (bf16 burst TFLOP/s)
  cpu_baseline : the oracle port of the same step (oracle/ref_steps.py, torch CPU fp32 = what the reference runs on a
              CPU host) timed on this box's cores on a bounded sample (B=8)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

VOX = 32 ** 3
FLOPS_D2_FWD = 2 * 17179.9e6          # Conv3d(64->128,k4,s2,p1) forward at B=64: SURVEY.md App. A.2 (dense MACs x 2)
STEP_GFLOP = {'wgan': 694.5, 'wgan_gp': 694.5 + 214.7}      # SURVEY.md 8d, as written, B=64


def workload_name(wl, b):
    if wl == 'autodecoder':
        return 'configs[2]: DeepSDF autodecoder 16384 pts x 512 shapes/step (train_sdf_autodecoder.py:84-91)'
    return 'configs[1]: 3D-CNN WGAN%s G+D step 32^3, batch %d/GPU (train_wgan.py:62-71 + :75-84)' % ('-GP' if wl == 'wgan_gp' else ' (clip)', b)


def peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d, 'measured (MEASURED_PEAKS.json)'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback (B200_PROFILING.md)'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []
        q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        self.cmd = ['nvidia-smi', '-i', str(index), '--query-gpu=' + q, '--format=csv,noheader,nounits', '-lms', '25']
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(self.cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([c.strip() for c in line.split(',')])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx or None, 'reasons': sorted(reasons), 'samples': len(sm)}


def synth_voxels(b, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.clamp(torch.randn((b, 32, 32, 32), generator=g) * 0.05, -0.1, 0.1) / 0.1      # SURVEY 8d / datasets.py:20-22


def dist_setup(n):
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1:
        os.environ['NCCL_DEBUG'] = 'WARN'          # keep NCCL's banner off stdout: rank 0 prints exactly one JSON line
        import torch.distributed as dist
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        return dist.get_rank(), world, local
    return 0, 1, 0


def max_over_ranks(ms, world):
    if world == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


# --------------------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the step (oracle port), on all host threads, bounded sample."""
    if rank != 0:
        return
    from oracle import ref_steps as S
    from oracle import shapes as TS
    b = 8
    wl = args.workload
    if wl.startswith('wgan'):
        gen = S.make_params(TS.gen_shapes(), 601)
        cri = S.make_params(TS.disc_shapes(), 602)
        step = S.WGANStepCPU(gen, cri, gp=(wl == 'wgan_gp'))
        real = synth_voxels(b, 605)
        z1 = torch.randn((b, 128), generator=torch.Generator().manual_seed(1))
        z2 = torch.randn((b, 128), generator=torch.Generator().manual_seed(2))
        alpha = torch.rand((b, 1, 1, 1), generator=torch.Generator().manual_seed(3))
        run = lambda: step(real, z1, z2, alpha)       # noqa: E731
        units, unit, metric = b * VOX, 'voxels/s', 'wgan_gd_step_voxels_per_s'
        sample = 'B=%d of the B=64 step, fp32, torch CPU' % b
    else:
        n, shapes = 65536, 8
        sd = S.make_params(TS.sdf_shapes(), 101)
        g = torch.Generator().manual_seed(5)
        pts = torch.rand((n, 3), generator=g) * 2 - 1
        table = torch.randn((shapes, 128), generator=g) * 0.01
        idx = (torch.arange(n) * shapes) // n
        sdf = torch.clamp(pts.norm(dim=1) - 0.5, -0.1, 0.1)
        step = S.AutodecoderStepCPU(sd, table)
        run = lambda: step(pts, sdf, idx)             # noqa: E731
        units, unit, metric = n, 'points/s', 'sdfnet_autodecoder_step_points_per_s'
        sample = '%d points x %d shapes of the 16384x512 step, fp32, torch CPU' % (n // shapes, shapes)
    cores, calib = pick_cpu_threads(run)
    for _ in range(max(1, min(args.warmup, 2))):
        run()
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / steps
    v = units / dt
    sample += '; threads picked by a one-step sweep %s' % calib
    print(json.dumps({
        'impl': 'reference', 'metric': metric, 'value': v, 'unit': unit, 'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup,
        'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args.workload, args.batch), 'global_batch': args.gpus * args.batch, 'parallelism': 'dp%d' % args.gpus,
                   'note': 'oracle port of the reference step (oracle/ref_steps.py = the reference modules\' torch-CPU fp32 path) on host cores, '
                           'bounded sample; the reference is pure PyTorch and /root/reference cannot travel to the GPU box'},
        'cpu_baseline': {'value': v, 'unit': unit, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def usable_cpus():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:                                    # cgroup v2 quota of the container, if any
        q, per = open('/sys/fs/cgroup/cpu.max').read().split()
        if q != 'max':
            n = max(1, min(n, int(float(q) / float(per) + 0.5)))
    except Exception:
        pass
    return n


def pick_cpu_threads(run):
    """torch CPU thread count that runs the reference step fastest on this host: os.cpu_count() threads oversubscribe
    the GPU boxes badly (128 logical CPUs, 22.9 s/step at 128 threads vs 1.5 s), so time one step at a few counts."""
    avail = usable_cpus()
    cands = sorted({c for c in (8, 16, 32, 64, avail) if c <= avail} | {min(avail, 8)})
    torch.set_num_threads(cands[0])
    run()                                   # warm-up (primitive caches, allocator)
    best, log = None, []
    for c in cands:
        torch.set_num_threads(c)
        t0 = time.perf_counter()
        run()
        dt = time.perf_counter() - t0
        log.append('%d:%.2fs' % (c, dt))
        if best is None or dt < best[1]:
            best = (c, dt)
    torch.set_num_threads(best[0])
    return best[0], '{' + ' '.join(log) + '}'


def cpu_baseline(workload):
    from oracle import ref_steps as S
    from oracle import shapes as TS
    b = 8
    gen = S.make_params(TS.gen_shapes(), 601)
    cri = S.make_params(TS.disc_shapes(), 602)
    step = S.WGANStepCPU(gen, cri, gp=(workload == 'wgan_gp'))
    real = synth_voxels(b, 605)
    z1 = torch.randn((b, 128), generator=torch.Generator().manual_seed(1))
    z2 = torch.randn((b, 128), generator=torch.Generator().manual_seed(2))
    alpha = torch.rand((b, 1, 1, 1), generator=torch.Generator().manual_seed(3))
    cores, calib = pick_cpu_threads(lambda: step(real, z1, z2, alpha))
    n = 3
    t0 = time.perf_counter()
    for _ in range(n):
        step(real, z1, z2, alpha)
    dt = (time.perf_counter() - t0) / n
    return {'value': b * VOX / dt, 'unit': 'voxels/s', 'cores': cores, 'kind': 'port',
            'sample': '%d steps at B=%d of the B=64 G+D step (oracle/ref_steps.py, torch CPU fp32, %d threads picked by a one-step sweep %s)' % (n, b, cores, calib)}


# --------------------------------------------------------------------------------------------------------- roofline probe
def roofline_probe(flush):
    """Time the dominant kernel alone: sg_igemm (MODE_CONV, halo-reuse variant) for Conv3d(64->128) at B=64, bf16."""
    from shapegan_b200 import _lib as L
    from shapegan_b200 import raw
    b, r, cin, cout = 64, 16, 64, 128
    x = torch.randn((1, b, r, r, r, cin), device='cuda').to(torch.bfloat16)
    w = torch.randn((cout, cin, 4, 4, 4), device='cuda') * 0.05
    img = raw.pack_conv_fwd(w, 1)
    y = torch.empty((1, b, r // 2, r // 2, r // 2, cout), dtype=torch.bfloat16, device='cuda')
    rows = b * (r // 2) ** 3

    def launch():
        raw.igemm(L.MODE_CONV, 1, x, (b, r, r, r, cin), rows, 64 * cin, img, cout, y, cout, act=L.ACT_LRELU)
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    ms = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    avg = sum(ms[2:-2]) / len(ms[2:-2])
    pk, src = peaks()
    achieved = FLOPS_D2_FWD / (avg * 1e-3) / 1e12
    return {'bound': 'tensor', 'achieved': achieved, 'peak': pk['bf16_tflops'], 'unit': 'TFLOP/s', 'frac': achieved / pk['bf16_tflops'],
            'traffic': 34.69e6, 'traffic_source': 'dram__bytes_read+write per launch, ncu --set full (profiles/r01h_ncu_prof_conv.txt); '
            'algorithmic DRAM bytes 33.6 MB input + 1.0 MB weights (the 8.4 MB output stays in L2)',
            'kernel': 'sg_igemm_halo_kernel MODE_CONV Conv3d(64->128,k4,s2,p1) fwd B=64 (34.36 GFLOP/launch)',
            'launch_ms': avg, 'peak_source': src + ', burst (kernel timed alone)',
            'note': 'halo-reuse variant (one strided TMA block serves 4 taps); measured per-SM ceilings of the plain one-tile-per-tap kernel '
                    '(tools/diag_conv.py, DESIGN.md 4): ~3.5 clk per gathered 128-byte TMA row, SS-mode 128x128x16 MMA ~108 clk'}


def sdfnet_probe(dev, world):
    """SDFNet Mpoints/s (BASELINE.json metric, second half): fused forward at the north_star's 250k-point batch and at configs[2]'s
    8.39 M points, plus the configs[2] autodecoder step.  Inputs exceed L2 at the large size; CUDA events, 3 warm-ups."""
    from model.sdf_net import SDFNet
    from shapegan_b200 import train
    torch.manual_seed(0)
    net = SDFNet()
    pk, src = peaks()
    out = {}

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    for tag, n, reps in (('fwd_250k_points', 250000, 20), ('fwd_8388608_points', 512 * 16384, 5)):
        pts = torch.rand((n, 3), device=dev) * 2 - 1
        table = torch.randn((512, 128), device=dev) * 0.1
        idx = ((torch.arange(n, device=dev) * 512) // n).to(torch.int32)
        with torch.no_grad():
            ms = timed(lambda: net(pts, table, idx), reps)
        tf = 0.921e6 * n / (ms * 1e-3) / 1e12
        out[tag] = {'mpoints_per_s': n / ms / 1e3, 'ms': ms, 'tflops': tf, 'frac_of_burst_peak': tf / pk['bf16_tflops']}
    n = 512 * 16384
    pts = torch.rand((n, 3), device=dev) * 2 - 1
    sdf = torch.clamp(pts.norm(dim=1) - 0.5, -0.1, 0.1)
    idx = (torch.arange(n, device=dev) // 16384).to(torch.int32)
    step = train.AutodecoderStep(net, torch.randn((512, 128), device=dev) * 0.01, world_size=1)
    ms = timed(lambda: step(pts, sdf, idx), 5)
    tf = 2.763e6 * n / (ms * 1e-3) / 1e12
    out['autodecoder_step_512x16384'] = {'mpoints_per_s': n / ms / 1e3, 'ms': ms, 'tflops': tf, 'frac_of_sustained_peak': tf / pk['bf16_tflops_sustained']}
    big = out['fwd_8388608_points']
    out['roofline'] = {'bound': 'tensor', 'achieved': big['tflops'], 'peak': pk['bf16_tflops'], 'unit': 'TFLOP/s', 'frac': big['frac_of_burst_peak'],
                       'traffic': 17.8e6 * 8, 'traffic_source': 'ncu dram bytes at 1 M points x 8 (profiles/r01g_ncu_prof_sdf.txt); algorithmic 16 B/point',
                       'kernel': 'sg_sdfnet_fwd_kernel, 0.921 MFLOP/point (SURVEY 8d)', 'peak_source': src + ', burst'}
    del step
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--workload', default='wgan', choices=['wgan', 'wgan_gp', 'autodecoder'])
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32x'])
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-sdfnet', action='store_true', help='skip the secondary SDFNet Mpoints/s block')
    ap.add_argument('--ad-shapes', type=int, default=512, help='autodecoder workload: number of shapes (16384 points each)')
    args = ap.parse_args()
    if args.impl == 'reference':
        rank = int(os.environ.get('RANK', '0'))
        run_reference(args, rank, int(os.environ.get('WORLD_SIZE', '1')))
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the B200 path has no CPU fallback; use --impl reference for the CPU arm)')
    args.warmup = max(args.warmup, 3)
    rank, world, local = dist_setup(args.gpus)
    dev = torch.device('cuda', local)
    from shapegan_b200 import _lib as L
    from shapegan_b200 import config, train
    config.set_precision(args.precision)
    lib = L.lib()

    if args.workload == 'autodecoder':
        return bench_autodecoder(args, rank, world, dev, lib)

    from model.gan import Discriminator, Generator
    torch.manual_seed(0)
    gen, cri = Generator(), Discriminator()
    gp = args.workload == 'wgan_gp'
    step = train.WGANStep(gen, cri, gp=gp, world_size=world)
    b = args.batch
    # synthetic inputs: pinned host staging (e2e) + static device buffers (graph inputs)
    h_real = synth_voxels(b, 605 + rank).pin_memory()
    h_z1 = torch.randn((b, 128), generator=torch.Generator().manual_seed(11 + rank)).pin_memory()
    h_z2 = torch.randn((b, 128), generator=torch.Generator().manual_seed(23 + rank)).pin_memory()
    h_alpha = torch.rand((b, 1, 1, 1), generator=torch.Generator().manual_seed(37 + rank)).pin_memory()
    h_loss = torch.zeros(2).pin_memory()
    d_real, d_z1, d_z2, d_alpha = (t.to(dev) for t in (h_real, h_z1, h_z2, h_alpha))
    d_loss = torch.zeros(2, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def body():
        cl, gl = step(d_real, d_z1, d_z2, d_alpha)
        d_loss[0].copy_(cl); d_loss[1].copy_(gl)

    launches0 = lib.sg_launch_count()
    for _ in range(args.warmup):
        body()
    torch.cuda.synchronize()
    launches_per_step = (lib.sg_launch_count() - launches0) // args.warmup
    graph, graph_note = None, 'eager'
    if not args.no_graph:
        try:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                body()
            torch.cuda.current_stream().wait_stream(s)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                body()
            graph_note = 'cuda_graph(whole G+D step: fwd+bwd+allreduce+optimizer)'
            for _ in range(2):
                graph.replay()
            torch.cuda.synchronize()
        except Exception as e:       # capture is an optimisation; the eager path is the same kernels
            graph, graph_note = None, 'eager (graph capture failed: %s)' % str(e).split('\n')[0][:120]
            torch.cuda.synchronize()
    run = graph.replay if graph is not None else body

    # ---------------------------------------------------------------- value: inputs resident, per-iteration events, L2 flush
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier(world)
    total_ms = 0.0
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record()
        torch.cuda.synchronize()
        total_ms += e0.elapsed_time(e1)
    barrier(world)
    ms_step = max_over_ranks(total_ms / args.steps, world)
    # ---------------------------------------------------------------- e2e: H2D of the step's inputs + D2H of the losses, every step
    barrier(world)
    e2e_ms = 0.0
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d_real.copy_(h_real, non_blocking=True); d_z1.copy_(h_z1, non_blocking=True); d_z2.copy_(h_z2, non_blocking=True)
        if gp:
            d_alpha.copy_(h_alpha, non_blocking=True)
        run()
        h_loss.copy_(d_loss, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        e2e_ms += e0.elapsed_time(e1)
    barrier(world)
    e2e_step = max_over_ranks(e2e_ms / args.steps, world)
    clocks = sampler.finish() if rank == 0 else None
    err = lib.sg_check_device_error()
    if rank != 0:
        return
    h2d = h_real.numel() * 4 + h_z1.numel() * 4 + h_z2.numel() * 4 + (h_alpha.numel() * 4 if gp else 0)
    roof = roofline_probe(flush)
    pk, _ = peaks()
    out = {
        'metric': 'wgan_gd_step_voxels_per_s', 'value': world * b * VOX / (ms_step * 1e-3), 'unit': 'voxels/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'bf16' if args.precision == 'bf16' else 'bf16x3(fp32x)', 'data': 'synthetic',
        'config': {'workload': workload_name(args.workload, b),
                   'global_batch': world * b, 'parallelism': 'dp%d' % world, 'launch': graph_note, 'l2': 'flushed (256 MiB memset) between timed iterations',
                   'step_gflop_as_written': STEP_GFLOP[args.workload] * b / 64.0,
                   'step_tflops': STEP_GFLOP[args.workload] * b / 64.0 / ms_step, 'step_frac_of_sustained_peak': STEP_GFLOP[args.workload] * b / 64.0 / ms_step / pk.get('bf16_tflops_sustained', 1400.0)},
        'e2e': {'value': world * b * VOX / (e2e_step * 1e-3), 'unit': 'voxels/s', 'ms_per_step': e2e_step, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 8},
        'gpu_launches': int(launches_per_step * args.steps * 2),
        'gpu_launches_per_step': int(launches_per_step),
        'roofline': roof, 'clocks': clocks, 'device_error_word': err,
        'losses': [float(h_loss[0]), float(h_loss[1])],
    }
    if not args.no_sdfnet:
        try:
            out['sdfnet'] = sdfnet_probe(dev, world)
        except Exception as e:            # secondary numbers must never cost the headline line
            out['sdfnet'] = {'error': str(e).split('\n')[0][:200]}
    if not args.no_cpu_baseline:
        out['cpu_baseline'] = cpu_baseline(args.workload)
    print(json.dumps(out))


def bench_autodecoder(args, rank, world, dev, lib):
    """BASELINE.json configs[2]: DeepSDF autodecoder 16384 pts x 512 shapes / step (train_sdf_autodecoder.py:84-91)."""
    from model.sdf_net import SDFNet
    from shapegan_b200 import train
    torch.manual_seed(0)
    shapes, per = args.ad_shapes, 16384
    n = shapes * per
    net = SDFNet()
    g = torch.Generator().manual_seed(5 + rank)
    pts = (torch.rand((n, 3), generator=g) * 2 - 1).to(dev)
    sdf = torch.clamp(pts.norm(dim=1) - 0.5, -0.1, 0.1)
    idx = (torch.arange(n, device=dev) // per).to(torch.int32)
    table = torch.randn((shapes, 128), generator=g).to(dev) * 0.01
    step = train.AutodecoderStep(net, table, world_size=world)
    for _ in range(max(args.warmup, 3)):
        step(pts, sdf, idx)
    torch.cuda.synchronize()
    barrier(world)
    total = 0.0
    steps = min(args.steps, 10)
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); loss = step(pts, sdf, idx); e1.record()
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    ms = max_over_ranks(total / steps, world)
    if rank != 0:
        return
    pk, src = peaks()
    tflops = 2.763e6 * n / (ms * 1e-3) / 1e12
    print(json.dumps({
        'metric': 'sdfnet_autodecoder_step_points_per_s', 'value': world * n / (ms * 1e-3), 'unit': 'points/s', 'n_gpus': world, 'steps': steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': workload_name('autodecoder', 0), 'points_per_gpu': n,
                   'l2': 'inputs (%.1f GB) exceed L2' % (n * 20 / 1e9), 'parallelism': 'dp%d' % world},
        'roofline': {'bound': 'tensor', 'achieved': tflops, 'peak': pk['bf16_tflops_sustained'], 'unit': 'TFLOP/s', 'frac': tflops / pk['bf16_tflops_sustained'],
                     'traffic': None, 'note': 'whole step, 2.763 MFLOP/point fwd+bwd (SURVEY 8d); peak = ' + src + ' sustained'},
        'loss': float(loss)}))


if __name__ == '__main__':
    main()
