#!/usr/bin/env python
"""Benchmark of the ViP-NeRF volumetric render path (BASELINE.json metric: rays/sec, coarse+fine 64+128).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp16|bf16x3|fp32]

A "step" is one pass of the hot path over one 4096-ray batch (the reference's `chunk`) of the workload
BASELINE.json quotes the metric on: LLFF 'fern' camera (NDC), 64 coarse + 128 fine samples, visibility head
on, eval forward (retraw=False, sec_views_vis=False).  Synthetic rays from the real fern intrinsics,
random-init weights of the reference architecture.

Prints ONE JSON line (rank 0):
  value            whole-job rays/s with inputs resident in HBM (one fused kernel per step, CUDA events on the launching
                   stream, L2 flushed between steps, max over ranks)
  parity_check     the maps of the LAST TIMED launch against the CPU oracle on 512 strided rays (untimed); the run exits
                   non-zero when a gate fails (also for precision_modes, e2e result and the sharded == unsharded check)
  e2e              the same metric through the reference-facing plugin with pinned HOST inputs copied in and the rendered
                   maps copied back inside the timed region (N > 1: the maps of ALL ranks, gathered on rank 0 by the render
                   kernels' peer stores); pipelined and with a host sync per step; the per-tensor flow for comparison
  roofline         algorithmic FLOPs of the fused kernel / its measured duration against the measured BURST dense-bf16
                   peak (the kernel is timed alone); `sustained` = 65,536-ray launches back to back vs the sustained peak
  precision_modes  rays/s and parity of the fp16 and bf16x3 arithmetics of the same kernel on the same batch
  train            the 4096-ray training iteration (tensor-core chains, fused losses), ms per step and tensor roofline
  cpu_baseline     the reference's CPU path on this box's host cores on a bounded sample

--impl reference times the reference's own CPU implementation of the path (the UNMODIFIED reference staged under
oracle/_ref, else the oracle port), rank 0 only, on the same `config`.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RAYS_PER_STEP = 4096
FLOP_PER_RAY = 303_890_432          # 256 MLP evaluations x 1,187,072 FLOP
METRIC = 'rays/sec (coarse+fine, 64+128 samples) at 1/2/4/8 B200; PSNR vs ref'
WORKLOAD = "LLFF 'fern' 3 input views, 64+128 coarse/fine, visibility head on, 4096-ray batches"


def model_configs(precision, ndc=True):
    mlp = dict(num_samples=64, netdepth=8, netwidth=256, points_positional_encoding_degree=10,
               views_positional_encoding_degree=4, use_view_dirs=True, view_dependent_rgb=True, predict_visibility=True)
    return {'data_loader': {'ndc': ndc},
            'model': dict(name='VipNeRFFused01', coarse_mlp=dict(mlp), fine_mlp=dict(mlp, num_samples=128), chunk=4096,
                          lindisp=False, netchunk=16384, perturb=True, raw_noise_std=1.0, white_bkgd=False,
                          precision=precision)}


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return {'sustained': d['bf16_tflops_sustained'], 'burst': d['bf16_tflops'], 'source': 'measured'}
    return {'sustained': 1400.0, 'burst': 1590.0, 'source': 'fallback'}   # B200_PROFILING.md fallback


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period_s=0.01):
        self.sm, self.max_mhz, self.reasons, self.error, self.power = [], None, set(), None, []
        self._stop = threading.Event()
        try:
            import pynvml
            self.nvml = pynvml
            pynvml.nvmlInit()
            visible = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(visible.split(',')[index]) if visible and visible.split(',')[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception as e:   # noqa: BLE001
            self.error = repr(e)
            return
        self.period = period_s
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        n = self.nvml
        names = {'hw_slowdown': 'nvmlClocksThrottleReasonHwSlowdown',
                 'hw_thermal_slowdown': 'nvmlClocksThrottleReasonHwThermalSlowdown',
                 'sw_thermal_slowdown': 'nvmlClocksThrottleReasonSwThermalSlowdown',
                 'sw_power_cap': 'nvmlClocksThrottleReasonSwPowerCap'}
        while not self._stop.is_set():
            try:
                self.sm.append(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    self.power.append(n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
                except Exception:   # noqa: BLE001
                    pass
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for key, attr in names.items():
                    if mask & getattr(n, attr, 0):
                        self.reasons.add(key)
            except Exception as e:   # noqa: BLE001
                self.error = repr(e)
                return
            time.sleep(self.period)

    def stop(self):
        if self.error is not None and not self.sm:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'samples': 0, 'reasons': [f'nvml: {self.error}']}
        self._stop.set()
        self.thread.join(timeout=2)
        return {'sm_mhz': statistics.median(self.sm) if self.sm else None, 'sm_max_mhz': self.max_mhz,
                'sm_mhz_min': min(self.sm) if self.sm else None, 'power_w_max': max(self.power) if self.power else None,
                'samples': len(self.sm), 'reasons': sorted(self.reasons)}


def reference_cpu_render():
    """(render_fn, kind): the reference's own CPU implementation of the path - the UNMODIFIED reference model staged under
    oracle/_ref (oracle/build_ref.py; kind 'reference') when present, else the oracle port (kind 'port')."""
    from oracle import build_ref
    from oracle import vipnerf_oracle as O
    sd = O.synth_state_dict(0)
    if build_ref.ref_available():
        cfg = model_configs('bf16')
        cfg['model']['name'] = 'VipNeRF01'
        del cfg['model']['precision']
        model = build_ref.load_ref_get_model()(cfg, None)
        model.load_state_dict(sd)
        model.eval()
        return (lambda b: model(dict(b))), 'reference'
    return (lambda b: O.render(sd, b, ndc=True)), 'port'


def cpu_oracle_rate(n_rays, reps, threads):
    """(rays/s, kind) of the reference's CPU path on `n_rays` rays of the workload."""
    import torch
    from oracle import vipnerf_oracle as O
    torch.set_num_threads(threads)
    render, kind = reference_cpu_render()
    batch = O.make_rays('fern', n_rays, seed=2)
    best = float('inf')
    with torch.no_grad():
        render(O.make_rays('fern', min(256, n_rays), seed=2))   # warm-up
        for _ in range(reps):
            t0 = time.perf_counter()
            render(batch)
            best = min(best, time.perf_counter() - t0)
    return n_rays / best, kind


def bench_config(rays_per_step):
    """The workload description BOTH arms print (identical dict, so the driver's same_config holds); everything that
    is specific to one arm (precision, kernel, the CPU arm's bounded sample) goes into `details`."""
    return {'workload': WORKLOAD, 'rays_per_step': rays_per_step, 'samples': '64+128', 'ndc': True,
            'l2': 'GPU arm: flushed between timed iterations (256 MiB memset, untimed); CPU arm: n/a'}


# parity_check gates per arithmetic: key family -> (median, p99, max) of |out - ref| / max|ref|
# (tests/test_gpu_bench_shapes.py holds the same numbers and their derivation)
_GATES = {
    'bf16': {'map': (1e-4, 2e-3, 1e-2), 'depth': (3e-3, 3e-2, 1e-1)},
    'fp16': {'map': (2e-5, 4e-4, 2e-3), 'depth': (6e-4, 6e-3, 1e-1)},
    'bf16x3': {'map': (1e-5, 1e-4, 1e-4), 'depth': (1e-5, 1e-3, 1e-2)},
    'fp32': {'map': (1e-5, 1e-4, 1e-4), 'depth': (1e-5, 1e-3, 1e-2)},
}
PARITY_KEYS = ('rgb_fine', 'acc_fine', 'depth_fine', 'depth_ndc_fine', 'rgb_coarse', 'acc_coarse', 'depth_coarse')


def parity_check(out, host_batch, precision, n_sub=512):
    """Compares the maps of a launch (the LAST TIMED one) with the CPU oracle on a strided `n_sub`-ray subset of the
    same batch (rays are independent).  Untimed.  Returns {'ok', 'against', per key {'median','p99','max'}}."""
    import torch
    from oracle import vipnerf_oracle as O
    R = host_batch['rays_o'].shape[0]
    idx = torch.arange(0, R, max(1, R // n_sub))[:n_sub]
    sub = {k: v[idx] for k, v in host_batch.items()}
    with torch.no_grad():
        ref = O.render(O.synth_state_dict(0), sub, ndc=True)
    res, ok = {}, True
    for k in PARITY_KEYS:
        got = out[k].detach()[idx.to(out[k].device)].cpu().double()
        d = ((got - ref[k].double()).abs() / ref[k].abs().max().clamp_min(1e-30)).flatten()
        med, p99, mx = d.median().item(), torch.quantile(d, 0.99).item(), d.max().item()
        g = _GATES[precision]['depth' if k.startswith('depth') else 'map']
        good = bool(torch.isfinite(d).all()) and med <= g[0] and p99 <= g[1] and mx <= g[2]
        ok = ok and good
        res[k] = {'median': med, 'p99': p99, 'max': mx, 'ok': good}
    return {'ok': ok, 'against': f'CPU oracle (pinned to the unmodified reference) on {len(idx)} strided rays of the timed batch',
            'norm': '|out - ref| / max|ref| per key', 'precision': precision, 'keys': res}


def time_launches(fn, steps, flush, pre=None):
    """CUDA-event windows around `steps` calls of fn on the current stream, L2 flushed (untimed) before each."""
    import torch
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    last = None
    for i in range(steps):
        if flush is not None:
            flush.zero_()
        starts[i].record()
        last = fn()
        ends[i].record()
    return starts, ends, last


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores, rank 0 only.
    Drives the UNMODIFIED reference model staged under oracle/_ref (oracle/build_ref.py; `kind: reference`) when it is
    present, else the oracle port (`kind: port`); same weights, same ray batch, same config as the GPU arm."""
    if rank != 0:
        return
    import torch
    from oracle import vipnerf_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    render, kind = reference_cpu_render()
    with torch.no_grad():
        probe = O.make_rays('fern', 256, seed=2)
        render(probe)
        t0 = time.perf_counter()
        render(probe)
        rate = 256 / (time.perf_counter() - t0)
        budget_s = 90.0
        n = int(rate * budget_s / max(1, args.steps + args.warmup))
        n = max(256, min(args.rays_per_step, (n // 256) * 256))
        batch = O.make_rays('fern', n, seed=2)
        for _ in range(args.warmup):
            render(batch)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            render(batch)
        total = time.perf_counter() - t0
    value = n * args.steps / total
    sample = (f'{n} rays of the {args.rays_per_step}-ray batch per step, fp32 torch CPU, '
              + ('unmodified reference VipNeRF01 (oracle/_ref)' if kind == 'reference' else 'oracle port'))
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'rays/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': bench_config(args.rays_per_step),
            'details': {'host': 'cpu', 'threads': threads, 'rays_timed_per_step': n, 'kind': kind},
            'cpu_baseline': {'value': value, 'unit': 'rays/s', 'cores': threads, 'kind': kind, 'sample': sample},
            'e2e': {'value': value, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def run_frame(args, rank, world, local_rank):
    """--workload frame: BASELINE config 5 - the LLFF 504x378 full-frame render (190,512 rays per step) STRONGLY
    scaled over the ranks: every rank generates its pixel range on the device (vipnerf_generate_rays) and renders it
    with the fused kernel, whose ray warps store the per-ray maps straight into rank 0's arrays (sharding.PeerGather;
    fallback: one grouped NCCL send/receive); rank 0 post-processes on the device and copies the finished frame (uint8
    image + depth maps) to the host.  All of that is inside the timed region; the
    only per-step host input is the camera pose.  Prints one JSON line (informational: the default workload is the
    4096-ray batch BASELINE.json quotes the metric on)."""
    import numpy
    import torch
    import torch.distributed as dist
    from oracle import vipnerf_oracle as O
    from vipnerf_b200 import sharding
    from vipnerf_b200.DataPreprocessorFactory import get_data_preprocessor
    from vipnerf_b200.ModelFactory import get_model

    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)
    sc = O.SCENES['fern_half' if args.scene == 'fern' else 'dtu']   # BASELINE config 5 (LLFF 504x378) / 4 (DTU 400x300)
    h, w, f, ndc = sc['h'], sc['w'], sc['f'], sc['ndc']
    R = h * w
    cfg = model_configs(args.precision, ndc)
    cfg['data_loader']['data_preprocessor_name'] = 'DataPreprocessorFused01'
    cfg['device'] = [local_rank]
    mc = {'resolution': [h, w], 'intrinsic': [[f, 0.0, w / 2], [0.0, f, h / 2], [0.0, 0.0, 1.0]],
          'average_pose': numpy.eye(4).tolist(), 'translation_scale': 1, 'near': sc['near'], 'far': sc['far'],
          'near_ndc': 0.0, 'far_ndc': 1.0}
    dp = get_data_preprocessor(cfg, 'test', model_configs=mc)
    model = get_model(cfg, mc)
    model.load_state_dict(O.synth_state_dict(0))
    model = model.to(device).eval()
    poses = [numpy.concatenate([O._pose_from_seed(100 + i), [[0, 0, 0, 1]]], 0).astype(numpy.float32) for i in range(8)]
    lo, hi = sharding.shard_range(R, rank, world)
    keys = ('rgb_fine', 'depth_fine', 'depth_var_fine') + (('depth_ndc_fine', 'depth_var_ndc_fine') if ndc else ())

    peer, gather_mode = None, 'none (1 GPU)'
    if world > 1:
        try:
            if args.gather == 'nccl':
                raise RuntimeError('--gather nccl')
            peer = [sharding.PeerGather({k: ((3,) if k == 'rgb_fine' else ()) for k in keys}, R, device) for _ in range(2)]
            gather_mode = 'peer stores from the render kernel into rank 0 symmetric memory + 1 device barrier (no NCCL call)'
        except Exception as e:   # noqa: BLE001
            gather_mode = f'NCCL grouped send/recv into preallocated arrays ({type(e).__name__}: {str(e)[:80]})'

    # the cameras of the trajectory (4x4 pose algebra + a 3x3 inverse on the host, DataPreprocessorFused.camera) are
    # prepared once, like the reference's tester prepares its pose list before the render loop (Tester01.py:203-211)
    cameras = [dp.camera(pose, preprocess_pose=False) for pose in poses]

    def step(i, shard=True):
        if not shard:     # the whole frame on this rank alone (the sharded == unsharded check)
            out = model(dp.generate(cameras[i % len(cameras)]))
            return dp.postprocess({k: out[k] for k in keys}, '_fine')
        batch = dp.generate(cameras[i % len(cameras)], lo, hi - lo)
        if peer is not None:
            pg = peer[i & 1]
            model(batch, out=pg.local_outputs())
            maps = pg.finish()
        else:
            out = model(batch)
            maps = {k: out[k] for k in keys}
            if world > 1:
                maps = sharding.gather_outputs(maps, R, None, 0)
        if rank == 0:
            return dp.postprocess(maps, '_fine')     # device post-processing + the single D2H copy (synchronises)
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for i in range(args.warmup):
            step(i)
        barrier()
        sampler = ClockSampler(local_rank)
        t_ms = 0.0
        frame = None
        for i in range(args.steps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            frame = step(i)
            e.record()
            torch.cuda.synchronize()
            t_ms += s.elapsed_time(e)
        barrier()
        clocks = sampler.stop()
        # sharded == unsharded: rank 0 renders the same frame alone; the finished frames must be identical
        sharded_ok = None
        if world > 1:
            i_chk = args.steps + (args.steps & 1)
            sharded = step(i_chk)
            barrier()
            if rank == 0:
                alone = step(i_chk, shard=False)
                sharded_ok = all(numpy.array_equal(sharded[k], alone[k]) for k in alone)
    total = torch.tensor([t_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    total_ms = total.item()
    if rank == 0:
        value = R * args.steps / (total_ms * 1e-3)
        line = {'metric': METRIC, 'value': value, 'unit': 'rays/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
                'scaling': 'strong', 'vs_baseline': None, 'dtype': args.precision, 'data': 'synthetic',
                'config': {'workload': ('LLFF 504x378' if args.scene == 'fern' else 'DTU 400x300') +
                                       f' full-frame render ({R:,} rays per step), rays sharded over the GPUs, '
                                       'on-device ray generation, one gather (see gather), device post-processing, finished frame to host',
                           'rays_per_step': R, 'samples': '64+128', 'ndc': ndc, 'precision': args.precision,
                           'frames_per_s': args.steps / (total_ms * 1e-3), 'gather': gather_mode},
                'sharded_equals_unsharded': sharded_ok,
                'clocks': clocks,
                'e2e': {'value': value, 'unit': 'rays/s', 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': int(frame['image'].nbytes + sum(frame[k].nbytes for k in frame if k != 'image')),
                        'note': 'the timed region IS end to end: pose in (kernel argument), finished frame out'},
                'gpu_launches': args.steps * 3}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


TRAIN_FLOP_PER_POINT = 2 * (593_536 + 557_056 + 593_536)   # forward + backward-data chain + parameter gradients (MACs x 2)
# Algorithmic HBM bytes per sample point of the tensor-core training step at one secondary view (two view directions per
# point); every fp32 activation / gradient array is written once and read once per consumer, weights stay in L2
# (derivation: DESIGN.md section 4.6):
#   forward 22,852 (encodings 512, ten layer products 19,456, heads 2,844, compositing 40)
#   + compositing backward and heads backward 2,672 + backward-data chain 26,116
#   + parameter gradients 31,012 (nine tensor-core products 17,920, three FFMA products 3,840, heads 2,084, column sums 7,168)
TRAIN_TC_BYTES_PER_POINT = 22_852 + 2_672 + 26_116 + 31_012   # 82,652
# The fp16 mode (train_precision='fp16'; derivation: DESIGN.md section 4.6): every saved activation and chain gradient is
# an fp16 array, the ReLU masks are bits, the views branch is one product launch per view direction
#   forward 11,460 (encodings 384, eight trunk layers 8,192 + 256 of mask bits, feature 1,024, views 2 x 896, heads / compositing 68)
#   + compositing backward 100 + heads backward 1,312 + backward-data chain 9,220
#   + parameter gradients 12,068 (eleven tensor-core products 11,008, heads 1,060)
TRAIN_F16_BYTES_PER_POINT = 11_460 + 100 + 1_312 + 9_220 + 12_068   # 34,160


def make_train_step(device, R, rng, train_precision, seed, world=1, loss_path='fused', graph=False):
    """One training iteration of BASELINE config 3 through the plugin exactly as Trainer01.train_one_iter (:61-107)
    drives it: pinned host rays in, zero_grad, model(batch) in train mode, the four losses, backward, Adam step.
    graph=True: the same iteration captured once as a CUDA graph (vipnerf_b200.training.GraphedTrainStep) and replayed.
    Returns (step_fn, model, h2d_bytes)."""
    import torch
    from oracle import vipnerf_oracle as O
    from vipnerf_b200 import sharding
    from vipnerf_b200.ModelFactory import get_model
    V = 1
    cfg = model_configs('bf16', ndc=True)
    cfg['model']['rng'] = rng
    cfg['model']['train_precision'] = train_precision
    model = get_model(cfg, None)
    model.load_state_dict(O.synth_state_dict(0))
    model = model.to(device).train()
    opt = torch.optim.Adam(model.parameters(), lr=5e-4, betas=(0.9, 0.999), capturable=bool(graph))
    host = {k: v.pin_memory() for k, v in O.make_rays('re10k', R, seed=seed, n_sec_views=V).items()}
    sup_host = O.make_supervision('re10k', R, V)
    sup = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in sup_host.items()}
    # The reference's loss computer interface (LossComputer01.compute_losses) with the four losses fused into the CUDA
    # step (vipnerf_b200/LossComputerFused01.py): loss values from one kernel, their gradients formed inside the
    # compositing backward.  --loss-path torch evaluates the same losses with torch ops on the output tensors.
    from vipnerf_b200.LossComputerFused01 import LossComputer
    cfg['losses'] = [{'name': 'MSE01', 'weight': 1}, {'name': 'VisibilityLoss01', 'weight': 0.1},
                     {'name': 'VisibilityPriorLoss01', 'iter_weights': {'0': 0, '30000': 0.001}},
                     {'name': 'SparseDepthMSE01', 'weight': 0.1}]       # runs/training/train0012/Configs.json:69-89
    computer = LossComputer(cfg)

    if graph:
        if world > 1 or loss_path != 'fused' or rng != 'device':
            raise SystemExit('--graph: one GPU, the fused losses and --rng device')
        from vipnerf_b200.training import GraphedTrainStep
        example = dict(host)
        example.update(sup)
        graphed = GraphedTrainStep(model, computer, opt, example, device)

        def graphed_step():
            return graphed(host)        # batch -> pinned staging buffers, one graph launch; returns the device loss

        return graphed_step, model, graphed.h2d_bytes

    def step():
        batch = {k: v.to(device, non_blocking=True) for k, v in host.items()}
        batch.update(sup)
        opt.zero_grad(set_to_none=True)
        out = model(batch)
        losses = computer.compute_losses(batch, out) if loss_path == 'fused' else computer._compute_torch(batch, out, False)
        loss = losses['TotalLoss']
        loss.backward()
        if world > 1:
            sharding.allreduce_gradients(model, average=True)
        opt.step()
        return loss

    return step, model, sum(v.numel() * 4 for v in host.values())


def hbm_peak_gbs():
    ppath = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(ppath):
        with open(ppath) as f:
            return json.load(f).get('hbm_gbs', 6464.3)
    return 6464.3


def train_record(device, steps=5, warmup=3, R=4096):
    """`train` sub-record of the default bench line: the 4096-ray training iteration of BASELINE config 3 in the fp16
    tensor-core mode, captured as one CUDA graph (device-side random draws), ms per step, the fraction of the HBM
    roofline of the step's own algorithmic traffic and of the TENSOR roofline (3.66 TFLOP per iteration); the tf32 mode
    (eager launches, the r02 mid-round state) next to it."""
    import torch
    peaks = measured_peaks()

    def measure(train_precision, graph):
        step, _, h2d = make_train_step(device, R, 'device', train_precision, seed=2, graph=graph)
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        t_ms, loss_value = 0.0, None
        for _ in range(steps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            loss_value = step().item()          # the step's result comes back to the host
            e.record()
            torch.cuda.synchronize()
            t_ms += s.elapsed_time(e)
        return t_ms / steps, loss_value, h2d

    ms, loss_value, h2d = measure('fp16', True)
    ms_tf32, _, _ = measure('tf32', False)
    tflops = R * 256 * TRAIN_FLOP_PER_POINT / (ms * 1e-3) / 1e12
    gbs = R * 256 * TRAIN_F16_BYTES_PER_POINT / (ms * 1e-3) / 1e9
    return {'workload': 'RealEstate-10K camera, 1 secondary view, 4096-ray training iteration: pinned host rays in, train-mode '
                        'forward, the four ViP-NeRF losses, backward, Adam, loss value out',
            'train_precision': 'fp16 (tcgen05 kind::f16 chains + parameter gradients on fp16 saved activations / scaled fp16 '
                               'gradients; the tf32 mode\'s gradient accuracy)', 'rng': 'device',
            'launch': 'one CUDA graph per iteration (vipnerf_b200.training.GraphedTrainStep)',
            'ms_per_step': ms, 'value': R / (ms * 1e-3), 'unit': 'rays/s', 'steps': steps,
            'flop_per_step': R * 256 * TRAIN_FLOP_PER_POINT, 'achieved_tflops': tflops,
            'roofline': {'bound': 'hbm', 'achieved': gbs, 'peak': hbm_peak_gbs(), 'unit': 'GB/s', 'frac': gbs / hbm_peak_gbs(),
                         'bytes_per_point': TRAIN_F16_BYTES_PER_POINT,
                         'peak_kind': 'measured HBM copy bandwidth (MEASURED_PEAKS.json); algorithmic bytes of the whole step',
                         'tensor_frac': tflops / peaks['sustained'],
                         'tensor_peak_kind': 'dense fp16 = the measured sustained bf16 rate'},
            'tf32_eager_ms_per_step': ms_tf32,
            'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4, 'final_loss': loss_value}


def run_train(args, rank, world, local_rank):
    """--workload train: BASELINE config 3 - one TRAINING iteration of the RealEstate-10K setup (2 input views -> one
    secondary view, 2048 + 2048 rays, the four losses of the shipped configs) through the plugin exactly as
    Trainer01.train_one_iter (:61-107) drives it: zero_grad, model(batch) in train mode, the losses, backward,
    Adam step.  Weak scaling: every rank steps its own batch and the gradients are summed with one NCCL all-reduce per
    step (what torch.nn.DataParallel's reduce does in the reference).  Informational; the default workload is the eval
    render BASELINE.json quotes the metric on (its line carries a `train` sub-record)."""
    import torch
    import torch.distributed as dist
    from oracle import vipnerf_oracle as O

    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)
    R, V = args.rays_per_step, 1
    step, model, h2d = make_train_step(device, R, args.rng, args.train_precision, seed=2 + rank, world=world,
                                       loss_path=args.loss_path, graph=args.graph)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    t_ms = 0.0
    loss_value = None
    for _ in range(args.steps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        loss = step()
        loss_value = loss.item()           # the step's result comes back to the host (4 bytes)
        e.record()
        torch.cuda.synchronize()
        t_ms += s.elapsed_time(e)
    barrier()
    clocks = sampler.stop()
    total = torch.tensor([t_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    total_ms = total.item()
    if rank == 0:
        value = world * R * args.steps / (total_ms * 1e-3)
        sm_mhz = clocks.get('sm_mhz') or 1900.0
        peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12     # fp32 FFMA lanes x 2 FLOP x measured SM clock
        achieved = value / world * 256 * TRAIN_FLOP_PER_POINT / 1e12
        if args.train_precision in ('tf32', 'fp16'):
            # the tensor-core step is HBM-bound: activations and gradients make one round trip per consumer
            hbm_peak = hbm_peak_gbs()
            bpp = TRAIN_F16_BYTES_PER_POINT if args.train_precision == 'fp16' else TRAIN_TC_BYTES_PER_POINT
            gbs = value / world * 256 * bpp / 1e9
            roofline = {'bound': 'hbm', 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak,
                        'traffic': None,
                        'peak_kind': 'measured HBM copy bandwidth (MEASURED_PEAKS.json), whole step',
                        'bytes_per_ray': 256 * bpp,
                        'note': 'algorithmic bytes of all kernels of the step / step time (per-kernel figures: '
                                'profiles/r02_ncu_summary.md)',
                        'tensor_tflops_algorithmic': achieved}
        else:
            roofline = {'bound': 'fp32-ffma', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                        'frac': achieved / peak, 'traffic': None,
                        'peak_kind': '148 SMs x 128 FFMA lanes x 2 x measured SM clock (CUDA cores)',
                        'flop_per_ray': 256 * TRAIN_FLOP_PER_POINT}
        line = {'metric': 'training rays/s (forward + 4 losses + backward + Adam), 64+128 samples', 'value': value,
                'unit': 'rays/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': {'fp32': 'f32', 'tf32': 'tf32', 'fp16': 'f16'}[args.train_precision], 'data': 'synthetic',
                'config': {'workload': 'RealEstate-10K camera, 2 input views (1 secondary view), full ViP-NeRF visibility + '
                                       'sparse-depth losses, one training iteration per step',
                           'rays_per_step_per_gpu': R, 'samples': '64+128', 'ndc': True,
                           'rng': args.rng + (' (torch CPU generator in the reference\'s draw order, uploaded per step)'
                                              if args.rng == 'reference' else ' (torch CUDA generator, same distributions)'),
                           'train_precision': args.train_precision,
                           'launch': 'one CUDA graph per iteration' if args.graph else 'eager launches',
                           'kernels': {'fp32': 'fp32 CUDA-core training path (k_mlp_fp32<save>, k_composite_bwd, k_mlp_bwd_fp32, k_gemm_tn)',
                                       'tf32': 'tensor-core training path: k_linear_tc (forward and backward-data chains) + k_gemm_tn_tc '
                                               '(parameter gradients) on tcgen05 kind::tf32; encodings, heads, compositing and its backward fp32',
                                       'fp16': 'tensor-core training path on fp16 arrays: k_linear_tc (chains; the views branch and both heads in '
                                               'its epilogues, ReLU masks as bits) + k_gemm_tn_tc (parameter gradients) on tcgen05 kind::f16, '
                                               'scaled fp16 gradients; encodings, compositing and its backward fp32'}[args.train_precision],
                           'l2': 'working set (5.6 - 11 KB per sample point saved, > 10 GB per step) exceeds L2 by construction',
                           'final_loss': loss_value},
                'clocks': clocks,
                'e2e': {'value': value, 'unit': 'rays/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                        'note': 'the timed region IS end to end: pinned host rays in, loss value out'},
                'gpu_launches': args.steps * {'tf32': 89, 'fp16': 91, 'fp32': 95}[args.train_precision],   # counted in the ncu launch lists (profiles/)
                'roofline': roofline}
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            torch.set_num_threads(threads)
            n_cpu = 128
            rays = O.make_rays('re10k', n_cpu, seed=2, n_sec_views=V)
            sup_c = O.make_supervision('re10k', n_cpu, V)
            sd = {k: v.clone().requires_grad_(True) for k, v in O.synth_state_dict(0).items()}
            best = float('inf')
            for _ in range(2):
                t0 = time.perf_counter()
                draws = O.draw_training_randoms(n_cpu)
                out = O.render(sd, rays, ndc=True, train_randoms=draws)
                mn, md = sup_c['indices_mask_nerf'], sup_c['indices_mask_sparse_depth']
                l = sum(torch.mean(torch.square(out[f'rgb_{t}'][mn] - sup_c['target_rgb'][mn])) for t in ('coarse', 'fine'))
                l = l + 0.1 * torch.mean(torch.square(out['depth_fine'][md] - sup_c['sparse_depth_values'][:, 0][md]))
                l.backward()
                best = min(best, time.perf_counter() - t0)
            line['cpu_baseline'] = {'value': n_cpu / best, 'unit': 'rays/s', 'cores': threads, 'kind': 'port',
                                    'sample': f'{n_cpu}-ray training step (forward + backward, torch autograd over the CPU oracle), best of 2'}
        else:
            line['cpu_baseline'] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp16', 'bf16x3', 'fp32'])
    ap.add_argument('--rays-per-step', type=int, default=RAYS_PER_STEP)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='batch', choices=['batch', 'frame', 'train'])
    ap.add_argument('--train-precision', default='fp32', choices=['fp32', 'tf32', 'fp16'],
                    help='train workload: tf32 = every 256-wide product of the step on the tensor cores (tcgen05 kind::tf32)')
    ap.add_argument('--rng', default='reference', choices=['reference', 'device'],
                    help='train workload: where the stratified jitter / cdf samples / density noise are drawn')
    ap.add_argument('--loss-path', default='fused', choices=['fused', 'torch'], help='train workload: fused CUDA losses or torch ops')
    ap.add_argument('--graph', action='store_true', help='train workload: the iteration captured once as a CUDA graph and replayed')
    ap.add_argument('--scene', default='fern', choices=['fern', 'dtu'], help='camera of the frame workload')
    ap.add_argument('--gather', default='peer', choices=['peer', 'nccl'], help='N > 1: how the rendered maps reach rank 0')
    ap.add_argument('--quick', action='store_true', help='batch workload: skip the sustained / precision_modes / train sub-records')
    ap.add_argument('--no-train', action='store_true', help='batch workload: skip the train sub-record')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.gpus > 1 and world == 1:
        # not under torchrun: launch one process per GPU ourselves
        port = 29500 + os.getpid() % 2000
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}',
               '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if args.workload == 'frame':
        run_frame(args, rank, world, local_rank)
        return
    if args.workload == 'train':
        run_train(args, rank, world, local_rank)
        return

    run_batch(args, rank, world, local_rank)


def _measure_kernel(model, precision, dev_batch, steps, warmup, flush, barrier):
    """Device-timed launches of the hot path with inputs resident in HBM.  Returns (per-step ms, last output)."""
    import torch
    from vipnerf_b200 import renderpath
    device = dev_batch['rays_o'].device
    packed_c = model._packed_weights('coarse', precision, device)
    packed_f = model._packed_weights('fine', precision, device)
    eval_keys = renderpath.pass_keys(True, False, 0)

    def hot_path():
        return renderpath.render_rays(dev_batch, packed_c, packed_f, ndc=True, precision=precision, keys=eval_keys)

    with torch.no_grad():
        for _ in range(warmup):
            hot_path()
        barrier()
        starts, ends, last = time_launches(hot_path, steps, flush)
        barrier()
    return [s.elapsed_time(e) for s, e in zip(starts, ends)], last


def run_batch(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from oracle import vipnerf_oracle as O                      # only for synthetic weights / rays, parity_check and cpu_baseline
    from vipnerf_b200 import sharding
    from vipnerf_b200.ModelFactory import get_model

    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl ours needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)

    R = args.rays_per_step
    precision = args.precision
    sd = O.synth_state_dict(0)

    def make_model(prec):
        m = get_model(model_configs(prec), None)
        m.load_state_dict(sd)
        return m.to(device).eval()

    model = make_model(precision)
    # weak scaling: every rank renders its own 4096-ray batch of the frame (different pixels per rank)
    host_batch = {k: v.pin_memory() for k, v in O.make_rays('fern', R, seed=2 + rank).items()}
    dev_batch = {k: v.to(device) for k, v in host_batch.items()}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ------------------------------------------------------------------ value: kernel path, inputs in HBM
    sampler = ClockSampler(local_rank)
    step_ms, last_out = _measure_kernel(model, precision, dev_batch, args.steps, args.warmup, flush, barrier)
    clocks = sampler.stop()
    total_ms = max_over_ranks(sum(step_ms))
    value = world * R * args.steps / (total_ms * 1e-3)
    # parity of the LAST TIMED launch (untimed): every rank checks its own batch, any failure fails the run
    check = parity_check(last_out, host_batch, precision)
    all_ok = max_over_ranks(0.0 if check['ok'] else 1.0) == 0.0

    # ------------------------------------------------------------------ e2e: plugin call with host buffers
    # Every step: the step's rays leave PINNED HOST memory, the plugin renders them, the rendered maps (N > 1: the maps
    # of ALL ranks, gathered on rank 0) arrive in pinned host memory.  Host side: vipnerf_b200.hostio - one pinned
    # buffer per direction (one H2D, one D2H) and, on one GPU, the three operations replayed as one CUDA graph around
    # `model(batch, out=...)`.  The reference's own per-tensor flow (14 small copies around `model(batch)`) is timed
    # next to it as `plain_plugin_call`.
    from vipnerf_b200 import hostio
    out_keys = ('rgb_fine', 'depth_fine', 'depth_var_fine', 'depth_ndc_fine', 'depth_var_ndc_fine')
    shapes = {k: ((3,) if k == 'rgb_fine' else ()) for k in out_keys}
    # N > 1: the gather of the rendered pixels on rank 0.  Preferred: fused with the render - every rank's kernel stores
    # its finished maps straight into rank 0's arrays over NVLink (sharding.PeerGather, torch symmetric memory) and one
    # device-side barrier per step publishes them; two buffer sets alternate so that one barrier per step suffices.
    # Fallback (no symmetric memory): one grouped NCCL send/receive into preallocated arrays (sharding.gather_outputs).
    peer, gather_mode, graphed = None, 'none (1 GPU)', None
    inputs = hostio.FlatBuffers({k: tuple(v.shape) for k, v in host_batch.items()}, device)
    for k, v in host_batch.items():
        inputs.host[k].copy_(v)
    if world > 1:
        try:
            if args.gather == 'nccl':
                raise RuntimeError('--gather nccl')
            peer = [sharding.PeerGather(shapes, world * R, device) for _ in range(2)]
            gather_mode = 'peer stores from the render kernel into rank 0 symmetric memory + 1 device barrier (no NCCL call)'
            host_flat = torch.empty(peer[0].total, dtype=torch.float32).pin_memory() if rank == 0 else None
        except Exception as e:   # noqa: BLE001
            peer = None
            gather_mode = f'NCCL grouped send/recv into preallocated arrays ({type(e).__name__}: {str(e)[:80]})'
        if peer is None:
            host_out = {k: torch.empty((world * R,) + shapes[k], dtype=torch.float32).pin_memory() for k in out_keys}
        d2h = (peer[0].total * 4) if peer is not None else sum(v.numel() * 4 for v in host_out.values())
    else:
        graphed = hostio.GraphedRender(model, host_batch, out_keys, device=device)
        d2h = graphed.d2h_bytes
    h2d = inputs.nbytes
    step_no = [0]

    def e2e_step():
        if graphed is not None:
            return graphed()           # one cudaGraphLaunch: H2D, fused render, D2H
        batch = inputs.upload()
        if peer is not None:
            pg = peer[step_no[0] & 1]
            step_no[0] += 1
            model(batch, out=pg.local_outputs())
            maps = pg.finish()
            if maps is not None:       # rank 0 brings the WHOLE gathered result to the host: one copy
                host_flat.copy_(pg.buf, non_blocking=True)
                return {k: pg.view(host_flat, k, 0, world * R) for k in out_keys}, maps
            return None
        out = model(batch)
        maps = sharding.gather_outputs({k: out[k] for k in out_keys}, world * R, None, 0)
        if maps is not None:
            for k in out_keys:
                host_out[k].copy_(maps[k], non_blocking=True)
            return host_out, maps
        return None

    host_plain = {k: torch.empty((R,) + shapes[k], dtype=torch.float32).pin_memory() for k in out_keys}

    def plain_step():                  # the reference Tester's flow: per-tensor copies around model(batch)
        batch = {k: v.to(device, non_blocking=True) for k, v in host_batch.items()}
        out = model(batch)
        for k in out_keys:
            host_plain[k].copy_(out[k], non_blocking=True)

    def timed_pipelined(fn, n):
        starts, ends, _ = time_launches(fn, n, flush)
        barrier()
        return max_over_ranks(sum(s.elapsed_time(e) for s, e in zip(starts, ends))) * 1e-3

    def timed_synced(fn, n):
        total = 0.0
        for _ in range(n):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            total += time.perf_counter() - t0
        barrier()
        return max_over_ranks(total)

    e2e_steps = args.steps
    with torch.no_grad():
        for _ in range(3):
            e2e_step()
            plain_step()
        barrier()
        # (a) pipelined: steps enqueued back to back like a serving loop (no host sync per step); every step's window -
        # H2D, render, (gather), D2H - is timed with its own event pair on the stream, L2 flush outside the windows
        e2e_value = world * R * e2e_steps / timed_pipelined(e2e_step, e2e_steps)
        # (b) synchronous: the host waits for every step's result before it issues the next one (wall clock per step)
        e2e_synced = world * R * e2e_steps / timed_synced(e2e_step, e2e_steps)
        plain_value = world * R * e2e_steps / timed_pipelined(plain_step, e2e_steps)
        plain_synced = world * R * e2e_steps / timed_synced(plain_step, e2e_steps)
    # the e2e path delivered the right pixels: its host result against the device result of the value path (N = 1)
    e2e_ok = True
    if graphed is not None:
        host_maps = graphed()
        graphed.synchronize()
        e2e_ok = all(torch.equal(host_maps[k], last_out[k].cpu()) for k in out_keys)

    # sharded == unsharded (N > 1): rank 0 re-renders every rank's batch alone and compares with what the gather
    # delivered - bit for bit (untimed)
    sharded_ok = None
    if world > 1:
        with torch.no_grad():
            res = e2e_step()
            barrier()
            if rank == 0:
                on_host, gathered = res
                sharded_ok = True
                for r in range(world):
                    b = {k: v.to(device) for k, v in O.make_rays('fern', R, seed=2 + r).items()}
                    alone = model(b)
                    for k in out_keys:
                        sharded_ok = sharded_ok and bool(torch.equal(alone[k], gathered[k][r * R:(r + 1) * R]))
                        sharded_ok = sharded_ok and bool(torch.equal(on_host[k][r * R:(r + 1) * R], alone[k].cpu()))

    # ------------------------------------------------------------------ sub-records (rank 0 of a 1-GPU run only)
    extras = {}
    if world == 1 and not args.quick:
        peaks = measured_peaks()
        # sustained: 65,536-ray launches back to back for ~0.5 s, against the SUSTAINED dense-bf16 peak
        big = {k: v.to(device) for k, v in O.make_rays('fern', 65536, seed=2).items()}
        s_sampler = ClockSampler(local_rank)
        s_ms, _ = _measure_kernel(model, precision, big, 30, 3, None, barrier)
        s_clocks = s_sampler.stop()
        s_rate = 65536 * len(s_ms) / (sum(s_ms) * 1e-3)
        extras['sustained'] = {'rays_per_launch': 65536, 'launches': len(s_ms), 'value': s_rate, 'unit': 'rays/s',
                               'achieved': s_rate * FLOP_PER_RAY / 1e12, 'peak': peaks['sustained'],
                               'frac': s_rate * FLOP_PER_RAY / 1e12 / peaks['sustained'],
                               'peak_kind': f"dense bf16 sustained, {peaks['source']}", 'clocks': s_clocks,
                               'ms_per_launch_first_last': [round(x, 3) for x in (s_ms[:3] + s_ms[-3:])],
                               'l2': 'not flushed: launches back to back'}
        del big
        # the other arithmetics of the same kernel on the same batch: throughput and parity of their last timed launch
        modes = {}
        for prec in ('fp16', 'bf16x3', 'bf16'):
            if prec == precision:
                continue
            m = make_model(prec)
            ms, out_m = _measure_kernel(m, prec, dev_batch, max(5, args.steps // 2), 3, flush, barrier)
            modes[prec] = {'value': R * len(ms) / (sum(ms) * 1e-3), 'unit': 'rays/s', 'ms_per_step': statistics.mean(ms),
                           'parity_check': parity_check(out_m, host_batch, prec)}
            all_ok = all_ok and modes[prec]['parity_check']['ok']
        extras['precision_modes'] = modes
        if not args.no_train:
            extras['train'] = train_record(device, steps=5, warmup=3)

    if rank == 0:
        peaks = measured_peaks()
        kernel_ms = statistics.mean(step_ms)
        achieved = R * FLOP_PER_RAY / (kernel_ms * 1e-3) / 1e12
        prof = {}
        ppath = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
        if not os.path.isfile(ppath):
            ppath = os.path.join(ROOT, 'profiles', 'r01_traffic.json')
        if os.path.isfile(ppath):
            with open(ppath) as f:
                prof = json.load(f)
        traffic = prof.get(precision)
        line = {
            'metric': METRIC, 'value': value, 'unit': 'rays/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': {'bf16': 'bf16', 'bf16x3': 'bf16x3', 'fp32': 'f32', 'fp16': 'f16'}[precision],
            'data': 'synthetic',
            'config': bench_config(R),
            'details': {'precision': precision, 'rays_per_step_per_gpu': R,
                        'kernel': 'k_render_tc (fused coarse+fine, 1 launch/step)' if precision != 'fp32'
                        else 'staged fp32 kernels (5 launches/step)',
                        'weights': 'random-init reference architecture, density head rescaled (oracle.synth_state_dict)'},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': 'rays/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'steps': e2e_steps,
                    'api': 'VipNeRFFused.forward(batch, out=...) via ModelFactory.get_model, host I/O through vipnerf_b200.hostio '
                           + ('(GraphedRender: H2D + render + D2H as one CUDA graph)' if world == 1 else
                              '(FlatBuffers: one H2D; one D2H of the gathered result on rank 0)'),
                    'timing': 'per-step CUDA-event windows (H2D + render (+ gather) + D2H of the whole result), '
                              'steps enqueued back to back',
                    'gather': gather_mode,
                    'value_host_sync_per_step': e2e_synced,
                    'timing_host_sync': 'wall clock per step with torch.cuda.synchronize() on both sides',
                    'plain_plugin_call': {'value': plain_value, 'value_host_sync_per_step': plain_synced,
                                          'what': 'model(batch) with the reference Tester\'s per-tensor copies (9 H2D + 5 D2H), '
                                                  'local maps only'},
                    'result_equals_value_path': e2e_ok},
            'gpu_launches': args.steps * (1 if precision != 'fp32' else 5),
            'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peaks['burst'], 'unit': 'TFLOP/s',
                         'frac': achieved / peaks['burst'], 'traffic': traffic,
                         'peak_kind': f"dense bf16 BURST (cuBLAS best of 10), {peaks['source']}: the kernel is timed alone, "
                                      '~0.9 ms launches with an L2 flush in between',
                         'peak_sustained': peaks['sustained'], 'frac_of_sustained': achieved / peaks['sustained'],
                         'flop_per_ray': FLOP_PER_RAY, 'kernel_ms': kernel_ms,
                         'tensor_pipe_active_ncu': prof.get('tensor_pipe_active'),
                         'sustained': extras.get('sustained')},
            'parity_check': check,
        }
        all_ok = all_ok and e2e_ok
        if sharded_ok is not None:
            line['sharded_equals_unsharded'] = sharded_ok
            all_ok = all_ok and sharded_ok
        if 'precision_modes' in extras:
            line['precision_modes'] = extras['precision_modes']
        if 'train' in extras:
            line['train'] = extras['train']
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            n_cpu = 1024
            rate, kind = cpu_oracle_rate(n_cpu, 3, threads)
            line['cpu_baseline'] = {'value': rate, 'unit': 'rays/s', 'cores': threads, 'kind': kind,
                                    'sample': f'{n_cpu} rays of the same workload, best of 3, fp32 torch CPU, '
                                              + ('unmodified reference VipNeRF01 (oracle/_ref)' if kind == 'reference' else 'oracle port')}
        else:
            line['cpu_baseline'] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not all_ok:
        raise SystemExit('bench.py: parity_check FAILED - the timed launch does not match the oracle (see the JSON line)')


if __name__ == '__main__':
    main()
