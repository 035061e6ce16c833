#!/usr/bin/env python
"""Benchmark of the HRFuser-T fusion-backbone hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one backbone forward over one batch of synthetic frames
(configs[1]: HRFuser-T nuScenes r640, camera + lidar + radar, 8 frames / GPU, bf16
mode).  Prints ONE JSON line (rank 0).

  value      frames/s, whole job, inputs resident in HBM, one CUDA graph replay per step
  e2e        frames/s through the public API with HOST (pinned) inputs: H2D of the
             three fp32 input tensors and D2H of the four fp32 feature maps inside
             the timed region (three pipelined slots)
  roofline   dominant hrfuser_b200 kernel group: algorithmic bytes per launch
              / mean launch duration measured with CUDA events
             around every C-ABI call of one un-graphed step, vs MEASURED_PEAKS.json
  cpu_baseline   the oracle port (oracle/hrfuser_oracle.py, fp32 torch CPU) on the
             box's host cores, bounded sample -- reported, not the target
`--impl reference` times that CPU port alone (the reference itself is Python +
mmcv and cannot travel to the GPU box).
"""
import argparse
import copy
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'hrfuser_t_backbone_frames_per_s'
L2_BYTES = 126e6


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d['hbm_gbs'], bf16_tflops=d['bf16_tflops'],
                    bf16_tflops_sustained=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0,
                source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                 '-lms', '100', '-i', str(self.gpu)], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                                'sw_power_cap'), r[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=mx,
                    samples=len(sm), reasons=sorted(reasons))


def build_net(workload, precision, device=None, seed=0):
    from hrfuser_b200 import WORKLOADS, HRFuserHRFormerBased, backbone_cfg
    from hrfuser_b200.utils import randomize_parameters
    variant, dataset, H, W = WORKLOADS[workload]
    cfg = backbone_cfg(variant, dataset)
    c = copy.deepcopy(cfg)
    c.pop('type')
    net = HRFuserHRFormerBased(**c, precision=precision)
    randomize_parameters(net, seed)
    net.eval()
    if device is not None:
        net.to(device)
    mod_ch = tuple(c.get('mod_in_channels', [3, 3]))
    return cfg, net, (H, W), mod_ch


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs use every host core."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def cpu_port_fps(workload, budget_s=12.0, max_iters=8):
    """frames/s of the CPU port (oracle) on a bounded sample: batch-1 forwards."""
    from hrfuser_b200.utils import synthetic_inputs
    use_all_host_threads()
    from oracle import hrfuser_oracle as O
    cfg, net, (H, W), mod_ch = build_net(workload, 'fp32')
    sd = net.state_dict()
    x, mods = synthetic_inputs(1, H, W, mod_ch, seed=0)
    with torch.no_grad():
        O.backbone_forward(sd, cfg, x, mods)                 # warm-up
        ts, t_end = [], time.perf_counter() + budget_s
        while len(ts) < max_iters and (time.perf_counter() < t_end or len(ts) < 3):
            t0 = time.perf_counter()
            O.backbone_forward(sd, cfg, x, mods)
            ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    return dict(value=1.0 / med, unit='frames/s', cores=torch.get_num_threads(), kind='port',
                sample=f'{len(ts)} batch-1 fp32 forwards of {workload} ({H}x{W}), median '
                       f'{med * 1e3:.0f} ms, oracle/hrfuser_oracle.py on {os.cpu_count()} host CPUs')


def run_reference(args):
    """--impl reference: the CPU implementation of the path, all host threads."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from hrfuser_b200.utils import synthetic_inputs
    from oracle import hrfuser_oracle as O
    use_all_host_threads()
    cfg, net, (H, W), mod_ch = build_net(args.workload, 'fp32')
    sd = net.state_dict()
    x, mods = synthetic_inputs(1, H, W, mod_ch, seed=0)
    steps = min(args.steps, 10)
    with torch.no_grad():
        for _ in range(min(args.warmup, 2)):
            O.backbone_forward(sd, cfg, x, mods)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.backbone_forward(sd, cfg, x, mods)
        dt = time.perf_counter() - t0
    fps = steps / dt
    sample = (f'each step = 1 frame (batch-1 fp32 forward, {H}x{W}) of the CPU port '
              f'oracle/hrfuser_oracle.py; {steps} steps')
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': min(args.warmup, 2),
        'ms_per_step': dt / steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': args.workload, 'frames_per_step': 1, 'modalities': len(mod_ch) + 1},
        'cpu_baseline': dict(value=fps, unit='frames/s', cores=torch.get_num_threads(),
                             kind='port', sample=sample),
        'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def run_ours(args):
    from hrfuser_b200 import dist as hdist
    from hrfuser_b200 import ops
    from hrfuser_b200.engine import GraphedForward
    from hrfuser_b200.utils import synthetic_inputs
    rank, world, local = hdist.init_from_env()
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    # one process per GPU, on the GPU's own NUMA node (pinned buffers are allocated below)
    # (N = 1 stays unbound: nothing competes for the host links, and the CPU baseline below
    # should see every host core)
    if world == 1 or os.environ.get('HRF_NUMA_BIND') == '0':
        numa = 'unbound (single rank)' if world == 1 else 'off (HRF_NUMA_BIND=0)'
    else:
        numa = hdist.bind_to_gpu_numa(local)
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)
    cfg, net, (H, Wd), mod_ch = build_net(args.workload, args.precision, dev, seed=0)
    engine = net.engine()

    # ---- inputs: R rotating sets so the per-step input read misses L2 ----------
    bytes_in = B * (3 + sum(mod_ch)) * H * Wd * 4
    R = max(2, int(L2_BYTES * 2 // bytes_in) + 1)
    host_sets = []
    for r in range(R):
        x, mods = synthetic_inputs(B, H, Wd, mod_ch, seed=100 * rank + r)
        host_sets.append([t.pin_memory() for t in (x, *mods)])
    dev_sets = [[t.to(dev) for t in hs] for hs in host_sets]

    # ---- one CUDA graph per input set -------------------------------------------
    # HRF_STEPS_IN_FLIGHT=n > 1 (experiment): n independent steps replay concurrently on n
    # streams (private memory pools); the default 1 runs the steps back to back.
    n_fly = max(1, int(os.environ.get('HRF_STEPS_IN_FLIGHT', '1')))
    graphs, pool = [], None
    for ds in dev_sets:
        g = GraphedForward(engine, ds[0], ds[1:], pool=None if n_fly > 1 else pool)
        pool = pool or g.pool()
        graphs.append(g)
    launches_per_step = graphs[0].launches
    fly_streams = [torch.cuda.Stream() for _ in range(n_fly)] if n_fly > 1 else None
    torch.cuda.synchronize()

    def run_steps(n):
        if fly_streams is None:
            for i in range(n):
                graphs[i % R]()
            return
        cur = torch.cuda.current_stream()
        for st in fly_streams:
            st.wait_stream(cur)
        for i in range(n):
            gi = i % R
            with torch.cuda.stream(fly_streams[gi % n_fly]):
                graphs[gi]()
        for st in fly_streams:
            cur.wait_stream(st)

    # ---- (a) device-resident throughput -----------------------------------------
    run_steps(Wm)
    torch.cuda.synchronize()
    hdist.barrier()
    sampler = ClockSampler(local).start() if rank == 0 else None
    t_wall = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_steps(K)
    e1.record()
    torch.cuda.synchronize()
    hdist.barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    dev_ms = hdist.max_over_ranks(e0.elapsed_time(e1), dev)
    clocks = sampler.stop() if sampler else None

    # ---- (b) end to end: pinned host inputs -> H2D -> forward -> D2H ----------
    n_slots = int(os.environ.get('HRF_E2E_SLOTS', '3'))   # pipelined steps in flight (2: 2189, 3: 2480, 4: 2449 frames/s)
    slots = []
    for s in range(n_slots):
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            xin = [torch.empty_like(t) for t in dev_sets[0]]
            # own memory pool: the slots replay concurrently on different streams
            g = GraphedForward(engine, xin[0], xin[1:], pool=None)
            outs_host = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in g.out]
        slots.append((st, xin, g, outs_host))
    torch.cuda.synchronize()
    d2h_bytes = sum(o.numel() * o.element_size() for o in slots[0][3])

    def e2e_step(i):
        st, xin, g, outs_host = slots[i % n_slots]
        with torch.cuda.stream(st):
            for d, h in zip(xin, host_sets[i % R]):
                d.copy_(h, non_blocking=True)
            g()
            for h, d in zip(outs_host, g.out):
                h.copy_(d, non_blocking=True)
    for i in range(Wm):
        e2e_step(i)
    torch.cuda.synchronize()
    hdist.barrier()
    t0 = time.perf_counter()
    for i in range(K):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_ms = hdist.max_over_ranks((time.perf_counter() - t0) * 1e3, dev)
    hdist.barrier()

    # ---- (c) per-kernel roofline: CUDA events around every C-ABI call ----------
    kernels, roof = {}, None
    if rank == 0:
        pk = peaks()
        was_concurrent, engine.concurrent = engine.concurrent, False   # isolate every kernel
        was_pdl = ops.set_pdl(False)        # ... including from its successor's overlapped prologue
        with torch.no_grad():
            engine.forward(dev_sets[0][0], dev_sets[0][1:])
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with ops.record() as rec:
                # park the stream behind a ~75 ms spin kernel so that every launch of the
                # step is already queued when the GPU reaches it: the per-call events then
                # time kernels, not the host's launch latency
                torch.cuda._sleep(150_000_000)
                ev0.record()
                engine.forward(dev_sets[1 % R][0], dev_sets[1 % R][1:])
                ev1.record()
            torch.cuda.synchronize()
        engine.concurrent = was_concurrent
        ops.set_pdl(was_pdl)
        eager_ms = ev0.elapsed_time(ev1)
        summ = rec.summary()
        ours_ms = sum(g['total_ms'] for g in summ.values())
        ridge = pk['bf16_tflops'] * 1e12 / (pk['hbm_gbs'] * 1e9)
        for (kind, Cc), g in sorted(summ.items(), key=lambda kv: -kv[1]['total_ms']):
            ai = g['flops'] / g['bytes']
            gbs = g['bytes'] / g['total_ms'] / 1e6
            tfs = g['flops'] / g['total_ms'] / 1e9
            bound = 'hbm' if ai < ridge else 'tensor'
            kernels[f'{kind}_c{Cc}'] = dict(
                calls=g['calls'], total_ms=round(g['total_ms'], 4), avg_ms=round(g['avg_ms'], 5),
                share_of_hrf_kernels=round(g['total_ms'] / ours_ms, 4), bound=bound,
                achieved_gbs=round(gbs, 2), achieved_tflops=round(tfs, 3),
                frac=round(gbs / pk['hbm_gbs'] if bound == 'hbm' else tfs / pk['bf16_tflops'], 5))
        top_key, top = next(iter(kernels.items()))
        (kind, Cc), g = max(summ.items(), key=lambda kv: kv[1]['total_ms'])
        per_launch_bytes = g['bytes'] / g['calls']
        roof = dict(kernel=top_key, bound=top['bound'],
                    achieved=top['achieved_gbs'] if top['bound'] == 'hbm' else top['achieved_tflops'],
                    peak=pk['hbm_gbs'] if top['bound'] == 'hbm' else pk['bf16_tflops'],
                    unit='GB/s' if top['bound'] == 'hbm' else 'TFLOP/s', frac=top['frac'],
                    traffic=None, peak_source=pk['source'],
                    algorithmic_bytes_per_call=per_launch_bytes, avg_call_ms=top['avg_ms'],
                    share_of_step=round(g['total_ms'] / eager_ms, 4),
                    hrf_kernels_share_of_step=round(ours_ms / eager_ms, 4),
                    eager_step_ms=round(eager_ms, 3),
                    note='per-kernel times from one serial (single-stream, un-graphed) step')
        tp = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.isfile(tp):
            roof['traffic'] = json.load(open(tp)).get(top_key)

    # ---- gather per-rank frame counts (the path's only collective) ------------
    frames = torch.tensor([float(B * K)], device=dev)
    total_frames = float(hdist.gather_frames(frames).sum())

    if rank == 0:
        cpu = cpu_port_fps(args.workload) if (world == 1 and not args.no_cpu_baseline) else None
        out = {
            'metric': METRIC, 'value': total_frames / (dev_ms / 1e3), 'unit': 'frames/s',
            'n_gpus': world, 'steps': K, 'warmup': Wm, 'ms_per_step': dev_ms / K,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
            'config': {'workload': args.workload, 'frames_per_gpu_per_step': B,
                       'input': f'{H}x{Wd}', 'modalities': len(mod_ch) + 1,
                       'precision_mode': args.precision, 'weights': 'random-init (seeded)',
                       'parallelism': f'scene-batch shards x{world}, no data-path collective',
                       'l2': f'inputs rotate over {R} sets ({R * bytes_in / 1e6:.0f} MB > 126 MB L2); '
                             'one CUDA graph per set',
                       'timing': 'CUDA events around K graph replays, max over ranks'},
            'e2e': {'value': total_frames / (e2e_ms / 1e3), 'unit': 'frames/s',
                    'h2d_bytes_per_step': bytes_in, 'd2h_bytes_per_step': d2h_bytes,
                    'ms_per_step': e2e_ms / K, 'host_affinity_rank0': numa,
                    'how': f'{n_slots} pipelined slots (stream + graph each): pinned fp32 host '
                           'inputs -> H2D -> forward -> D2H of the 4 fp32 maps; wall clock '
                           'around K steps incl. final sync, max over ranks'},
            'gpu_launches': launches_per_step * K,
            'hrf_kernel_launches_per_step': launches_per_step,
            'wall_ms_timed_region': wall_ms,
            'clocks': clocks, 'roofline': roof, 'kernels': kernels, 'cpu_baseline': cpu,
        }
        print(json.dumps(out))
    hdist.barrier()
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='hrfuser_t_nus_r640')
    ap.add_argument('--batch', type=int, default=8, help='frames per GPU per step')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if not torch.cuda.is_available():
        sys.exit('bench.py: no CUDA device; the product path has no CPU fallback '
                 '(use --impl reference for the CPU port)')
    run_ours(args)


if __name__ == '__main__':
    main()
