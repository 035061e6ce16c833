#!/usr/bin/env python
"""Benchmark of the HRFuser-T fusion-backbone hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one backbone forward over one batch of synthetic frames
(configs[1]: HRFuser-T nuScenes r640, camera + lidar + radar, 8 frames / GPU, bf16
mode).  Prints ONE JSON line (rank 0).

  value      frames/s, whole job, inputs resident in HBM, one CUDA graph replay per step
  e2e        frames/s through the public API, `HRFuserHRFormerBased.forward(x_host, mods_host)`,
             with HOST (pinned) fp32 inputs: H2D of the input tensors and D2H of the four fp32
             feature maps inside the timed region (the caller pipelines three streams; the
             forward keeps one captured CUDA graph per input signature and stream)
  e2e_raw    the same through `pipeline.RawBackbone`: the host ships RAW frames (uint8 camera,
             fp32 lidar / radar at the un-padded size) and Normalize / Pad / format run on the
             device inside the same graph (49.8 MB instead of 70.8 MB per step)
  gpu_eager_baseline   the reference module (or, when no reference is reachable, its torch
             mirror hrfuser_b200/modules.py) in PyTorch eager on the same GPU, fp32 and bf16
             autocast, same batch -- SURVEY.md 8(d)'s "real bar"
  roofline   dominant hrfuser_b200 kernel group: algorithmic bytes per launch
              / mean launch duration measured with CUDA events
             around every C-ABI call of one un-graphed step, vs MEASURED_PEAKS.json
  cpu_baseline   the UNMODIFIED reference module on the box's host cores when a reference tree
             is reachable ($HRFUSER_REF, /root/reference, baseline/_ref: kind "reference"), else
             the oracle port (oracle/hrfuser_oracle.py, kind "port"); bounded sample --
             reported, not the target
`--impl reference` times that CPU arm alone, one frame per step.
`--workload hrfuser_t_stf_r1248` (configs[2]) and `--workload hrfuser_b_nus_r640 --train`
(configs[3]: SyncBN training step, forward + backward + SGD) select the other configs.
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
# raw (un-padded) sensor frame size per workload (H, W): the reference pads to /32 on the way in
RAW_SIZES = {'hrfuser_t_nus_r640': (360, 640), 'hrfuser_b_nus_r640': (360, 640)}


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


def cpu_arm(workload):
    """(forward(x, mods) -> maps, kind, what): the reference's own CPU implementation of the path
    when a reference tree is reachable, else the oracle port."""
    cfg, net, (H, W), mod_ch = build_net(workload, 'fp32')
    sd = net.state_dict()
    try:
        from oracle import ref_loader
        root = ref_loader.find_reference()
        if root is not None:
            ref = ref_loader.build_reference_backbone(cfg, root)
            ref.load_state_dict(sd)
            return (lambda x, mods: ref(x, list(mods))), 'reference', \
                f'unmodified reference HRFuserHRFormerBased from {root}', (H, W), mod_ch
    except Exception as e:                       # noqa: BLE001 -- any import problem: use the port
        sys.stderr.write(f'bench.py: reference not loadable ({type(e).__name__}: {e}); using the port\n')
    from oracle import hrfuser_oracle as O
    return (lambda x, mods: O.backbone_forward(sd, cfg, x, list(mods))), 'port', \
        'oracle/hrfuser_oracle.py', (H, W), mod_ch


def cpu_baseline_fps(workload, budget_s=12.0, max_iters=8):
    """frames/s of the CPU arm on a bounded sample: batch-1 fp32 forwards."""
    from hrfuser_b200.utils import synthetic_inputs
    use_all_host_threads()
    fwd, kind, what, (H, W), mod_ch = cpu_arm(workload)
    x, mods = synthetic_inputs(1, H, W, mod_ch, seed=0)
    with torch.no_grad():
        fwd(x, mods)                                         # warm-up
        ts, t_end = [], time.perf_counter() + budget_s
        while len(ts) < max_iters and (time.perf_counter() < t_end or len(ts) < 3):
            t0 = time.perf_counter()
            fwd(x, mods)
            ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    return dict(value=1.0 / med, unit='frames/s', cores=torch.get_num_threads(), kind=kind,
                sample=f'{len(ts)} batch-1 fp32 forwards of {workload} ({H}x{W}), median '
                       f'{med * 1e3:.0f} ms, {what} on {os.cpu_count()} host CPUs')


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, all host threads."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from hrfuser_b200.utils import synthetic_inputs
    use_all_host_threads()
    fwd, kind, what, (H, W), mod_ch = cpu_arm(args.workload)
    x, mods = synthetic_inputs(1, H, W, mod_ch, seed=0)
    steps, warm = args.steps, min(max(args.warmup, 1), 3)
    budget = float(os.environ.get('HRF_REF_BUDGET_S', '240'))
    with torch.no_grad():
        for _ in range(warm):
            fwd(x, mods)
        t0, done = time.perf_counter(), 0
        for _ in range(steps):
            fwd(x, mods)
            done += 1
            if time.perf_counter() - t0 > budget:          # bounded: never past a few minutes
                break
        dt = time.perf_counter() - t0
    fps = done / dt
    sample = (f'each step = 1 frame (batch-1 fp32 forward, {H}x{W}) of {what}; {done} steps')
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s',
        'n_gpus': args.gpus, 'steps': done, 'warmup': warm,
        'ms_per_step': dt / done * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': args.workload, 'frames_per_step': 1, 'modalities': len(mod_ch) + 1},
        'cpu_baseline': dict(value=fps, unit='frames/s', cores=torch.get_num_threads(),
                             kind=kind, sample=sample),
        'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def gpu_eager_baseline(workload, B, dev, iters=5):
    """The reference's own execution model on the same GPU: the module in PyTorch eager,
    fp32 and bf16 autocast."""
    from hrfuser_b200.utils import synthetic_inputs
    cfg, net, (H, W), mod_ch = build_net(workload, 'fp32')
    impl, mod = 'mirror (hrfuser_b200/modules.py, torch eager)', None
    try:
        from oracle import ref_loader
        root = ref_loader.find_reference()
        if root is not None:
            mod = ref_loader.build_reference_backbone(cfg, root)
            mod.load_state_dict(net.state_dict())
            impl = f'unmodified reference module from {root}'
    except Exception:                            # noqa: BLE001
        mod = None
    fwd = (lambda x, m: mod(x, list(m))) if mod is not None else (lambda x, m: net._forward_autograd(x, list(m)))
    (mod if mod is not None else net).to(dev)
    x, mods = synthetic_inputs(B, H, W, mod_ch, seed=7)
    x, mods = x.to(dev), [m.to(dev) for m in mods]
    res = {}
    with torch.no_grad():
        for name, ctx in (('fp32', torch.autocast('cuda', enabled=False)),
                          ('bf16_autocast', torch.autocast('cuda', dtype=torch.bfloat16))):
            with ctx:
                for _ in range(2):
                    fwd(x, mods)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(iters):
                    fwd(x, mods)
                b.record()
                torch.cuda.synchronize()
            ms = a.elapsed_time(b) / iters
            res[name] = dict(ms_per_step=round(ms, 3), frames_per_s=round(B / ms * 1e3, 1))
    res['impl'], res['frames_per_step'] = impl, B
    return res


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
    # Two independent steps replay concurrently on two streams (private memory pools), as a serving
    # process drives the backbone (and as the e2e leg's three caller streams do): a step's stem phase
    # -- one persistent conv kernel in flight at a time -- overlaps the previous step's stage tail.
    # `value` is K steps over the time they take this way; `serial` in the line is the same K steps
    # back to back on one stream (HRF_STEPS_IN_FLIGHT=1 makes that the headline).
    n_fly = max(1, int(os.environ.get('HRF_STEPS_IN_FLIGHT', '2')))
    graphs, pool = [], None
    for ds in dev_sets:
        g = GraphedForward(engine, ds[0], ds[1:], pool=None if n_fly > 1 else pool)
        pool = pool or g.pool()
        graphs.append(g)
    launches_per_step = graphs[0].launches
    fly_streams = [torch.cuda.Stream() for _ in range(n_fly)] if n_fly > 1 else None
    torch.cuda.synchronize()

    def run_steps(n, serial=False):
        if fly_streams is None or serial:
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
    serial_ms = dev_ms
    if fly_streams is not None:                  # the same K steps back to back on one stream
        run_steps(Wm, serial=True)
        torch.cuda.synchronize()
        hdist.barrier()
        e0.record()
        run_steps(K, serial=True)
        e1.record()
        torch.cuda.synchronize()
        hdist.barrier()
        serial_ms = hdist.max_over_ranks(e0.elapsed_time(e1), dev)

    # ---- (b) end to end through the public API: pinned host inputs -> forward -> D2H ---------
    # `net(x_host, mods_host)`: the forward copies the pinned host tensors straight into the
    # static buffers of its captured graph (H2D), replays it and returns clones of the four
    # maps; the caller pipelines n_slots streams (one graph instance per stream) and reads the
    # maps back into pinned host memory.
    n_slots = int(os.environ.get('HRF_E2E_SLOTS', '3'))   # pipelined steps in flight (2: 2189, 3: 2480, 4: 2449 frames/s)
    slot_streams = [torch.cuda.Stream() for _ in range(n_slots)]
    out_shapes = [tuple(o.shape) for o in graphs[0].out]
    outs_host = [[torch.empty(sh, dtype=torch.float32).pin_memory() for sh in out_shapes]
                 for _ in range(n_slots)]
    d2h_bytes = sum(o.numel() * o.element_size() for o in outs_host[0])

    def time_e2e(step_fn):
        for i in range(max(Wm, 3 * n_slots)):             # >= 2 calls per stream: the second captures
            step_fn(i)
        torch.cuda.synchronize()
        hdist.barrier()
        t0 = time.perf_counter()
        for i in range(K):
            step_fn(i)
        torch.cuda.synchronize()
        ms = hdist.max_over_ranks((time.perf_counter() - t0) * 1e3, dev)
        hdist.barrier()
        return ms

    def e2e_step(i):
        with torch.cuda.stream(slot_streams[i % n_slots]), torch.no_grad():
            hs = host_sets[i % R]
            outs = net(hs[0], hs[1:])
            for h, d in zip(outs_host[i % n_slots], outs):
                h.copy_(d, non_blocking=True)
    e2e_ms = time_e2e(e2e_step)

    # ---- (b') raw-frame leg: uint8 camera + fp32 lidar / radar at the un-padded size ---------
    from hrfuser_b200.pipeline import Normalize, RawBackbone
    raw_hw = RAW_SIZES.get(args.workload, (H, Wd))
    keys = ['img'] + [f'mod{k}_img' for k in range(len(mod_ch))]
    norms = [Normalize([123.675, 116.28, 103.53], [58.395, 57.12, 57.375], to_rgb=True, keys=['img'])]
    norms += [Normalize([0.5] * c, [2.0] * c, to_rgb=False, keys=[keys[k + 1]], sensor_type='lidar')
              for k, c in enumerate(mod_ch)]
    rb = RawBackbone(net, norms)
    g = torch.Generator().manual_seed(5 + rank)
    raw_sets = []
    for r in range(R):
        d = {'img': torch.randint(0, 256, (B, *raw_hw, 3), generator=g, dtype=torch.uint8).pin_memory()}
        for k, c in enumerate(mod_ch):
            d[keys[k + 1]] = torch.randn(B, *raw_hw, c, generator=g).pin_memory()
        raw_sets.append(d)
    raw_bytes = sum(t.numel() * t.element_size() for t in raw_sets[0].values())

    def raw_step(i):
        with torch.cuda.stream(slot_streams[i % n_slots]):
            outs = rb(raw_sets[i % R])
            for h, d in zip(outs_host[i % n_slots], outs):
                h.copy_(d, non_blocking=True)
    raw_ms = time_e2e(raw_step)

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
                calls=g['calls'], tensors=g['problems'], total_ms=round(g['total_ms'], 4), avg_ms=round(g['avg_ms'], 5),
                share_of_hrf_kernels=round(g['total_ms'] / ours_ms, 4), bound=bound,
                achieved_gbs=round(gbs, 2), achieved_tflops=round(tfs, 3),
                frac=round(gbs / pk['hbm_gbs'] if bound == 'hbm' else tfs / pk['bf16_tflops'], 5))
        top_key, top = next(iter(kernels.items()))
        (kind, Cc), g = max(summ.items(), key=lambda kv: kv[1]['total_ms'])
        # grouped launches (the camera's branch 0 + the modality streams in one call) are counted per
        # PROBLEM: bytes per tensor, and the graph-timed call below is a one-tensor call
        per_launch_bytes = g['bytes'] / g['problems']
        roof = dict(kernel=top_key, bound=top['bound'],
                    achieved=top['achieved_gbs'] if top['bound'] == 'hbm' else top['achieved_tflops'],
                    peak=pk['hbm_gbs'] if top['bound'] == 'hbm' else pk['bf16_tflops'],
                    unit='GB/s' if top['bound'] == 'hbm' else 'TFLOP/s', frac=top['frac'],
                    traffic=None, peak_source=pk['source'],
                    algorithmic_bytes_per_call=per_launch_bytes, avg_call_ms=round(g['total_ms'] / g['problems'], 5),
                    launches_per_step=g['calls'], tensors_per_step=g['problems'],
                    share_of_step=round(g['total_ms'] / eager_ms, 4),
                    hrf_kernels_share_of_step=round(ours_ms / eager_ms, 4),
                    eager_step_ms=round(eager_ms, 3),
                    note='per-kernel times from one serial (single-stream, un-graphed) step')
        tp = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.isfile(tp):
            roof['traffic'] = json.load(open(tp)).get(top_key)
        # The event-bracketed time of a call includes its launch latency (every call sits between two
        # event records, nothing overlaps it).  Second measurement of the dominant kernel: the same
        # call (same tensors, same blob) captured 20 times into ONE CUDA graph and replayed -- what
        # the kernel costs back to back, as it runs inside the step's graph.
        try:
            g_ms = graph_time_call(engine, kind, Cc, dev_sets[0])
        except Exception as e:                       # noqa: BLE001 -- a diagnostic must not fail the run
            g_ms = None
            roof['graph_timing_error'] = f'{type(e).__name__}: {e}'
        if g_ms:
            roof['avg_call_ms_events'] = roof['avg_call_ms']
            roof['achieved_events'], roof['frac_events'] = roof['achieved'], roof['frac']
            roof['avg_call_ms'] = round(g_ms, 5)
            if top['bound'] == 'hbm':
                roof['achieved'] = round(per_launch_bytes / g_ms / 1e6, 2)
            else:
                roof['achieved'] = round(g['flops'] / g['problems'] / g_ms / 1e9, 3)
            roof['frac'] = round(roof['achieved'] / roof['peak'], 5)
            roof['note'] = ('avg_call_ms / achieved / frac: the dominant kernel replayed 20x back to back in one CUDA '
                            'graph, CUDA events around 5 replays (input L2-resident, as inside the step where its '
                            'producer has just written it); *_events: the same call bracketed by two event records '
                            'in a serial un-graphed step (includes its launch latency); share_of_step and the '
                            '`kernels` table use the event times; per TENSOR: in the step the camera and modality '
                            'streams share grouped launches (tensors_per_step / launches_per_step), '
                            'algorithmic_bytes_per_call and avg_call_ms are for one tensor')

    # ---- gather per-rank frame counts (the path's only collective) ------------
    frames = torch.tensor([float(B * K)], device=dev)
    total_frames = float(hdist.gather_frames(frames).sum())

    if rank == 0:
        cpu = cpu_baseline_fps(args.workload) if (world == 1 and not args.no_cpu_baseline) else None
        eager = None
        if world == 1 and not args.no_eager_baseline:
            try:
                eager = gpu_eager_baseline(args.workload, B, dev)
            except Exception as e:               # noqa: BLE001 -- a baseline must not fail the run
                eager = {'error': f'{type(e).__name__}: {e}'}
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
                       'steps_in_flight': n_fly,
                       'timing': f'CUDA events around K graph replays ({n_fly} independent steps in flight on '
                                 f'{n_fly} streams), max over ranks'},
            'serial': {'value': total_frames / (serial_ms / 1e3), 'ms_per_step': serial_ms / K,
                       'what': 'the same K steps back to back on one stream'},
            'e2e': {'value': total_frames / (e2e_ms / 1e3), 'unit': 'frames/s',
                    'h2d_bytes_per_step': bytes_in, 'd2h_bytes_per_step': d2h_bytes,
                    'ms_per_step': e2e_ms / K, 'host_affinity_rank0': numa,
                    'h2d_gb_per_s': round(bytes_in / (e2e_ms / K) / 1e6, 2),
                    'd2h_gb_per_s': round(d2h_bytes / (e2e_ms / K) / 1e6, 2),
                    'how': f'through HRFuserHRFormerBased.forward(x_host, mods_host) on {n_slots} '
                           'caller streams (the forward keeps one captured graph per signature and '
                           'stream): pinned fp32 host inputs -> H2D -> forward -> D2H of the 4 fp32 '
                           'maps; wall clock around K steps incl. final sync, max over ranks'},
            'e2e_raw': {'value': total_frames / (raw_ms / 1e3), 'unit': 'frames/s',
                        'h2d_bytes_per_step': raw_bytes, 'd2h_bytes_per_step': d2h_bytes,
                        'ms_per_step': raw_ms / K,
                        'how': f'through pipeline.RawBackbone on {n_slots} caller streams: pinned RAW '
                               f'host frames ({raw_hw[0]}x{raw_hw[1]}: uint8 camera, fp32 modalities) -> '
                               'H2D -> hrf_input_prologue_fwd + forward in one graph -> D2H of the maps'},
            'gpu_eager_baseline': eager,
            'gpu_launches': launches_per_step * K,
            'hrf_kernel_launches_per_step': launches_per_step,
            'wall_ms_timed_region': wall_ms,
            'clocks': clocks, 'roofline': roof, 'kernels': kernels, 'cpu_baseline': cpu,
        }
        print(json.dumps(out))
    hdist.barrier()
    if world > 1:
        torch.distributed.destroy_process_group()


def graph_time_call(engine, kind, Cc, inputs, n=20, reps=5):
    """ms per call of the first (kind, C) C-ABI call of an engine forward, replayed n times back to
    back inside one CUDA graph (None when the kind has no replayable wrapper)."""
    from hrfuser_b200 import ops
    fname = {'mixffn': 'mixffn', 'lsa': 'window_attention', 'mwca': 'window_attention'}.get(kind)
    if fname is None:
        return None
    orig, captured = getattr(ops, fname), []

    def spy(*a, **k):
        with ops.record() as rec:
            out = orig(*a, **k)
        if not captured and rec.calls and rec.calls[0][0] == kind and rec.calls[0][1].get('C') == Cc:
            captured.append((a, k))
        return out
    was = engine.concurrent
    engine.concurrent = False
    setattr(ops, fname, spy)
    try:
        with torch.no_grad():
            engine.forward(inputs[0], inputs[1:])
    finally:
        setattr(ops, fname, orig)
        engine.concurrent = was
    torch.cuda.synchronize()
    if not captured:
        return None
    a, k = captured[0]
    st = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.stream(st):
        orig(*a, **k)
        st.synchronize()
        with torch.cuda.graph(graph, stream=st):
            for _ in range(n):
                orig(*a, **k)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = float('inf')
    for _ in range(reps):
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best


def run_train(args):
    """configs[3]: HRFuser-B nuScenes SyncBN training step (forward + backward + SGD) with the
    scene batch sharded over the ranks; the SyncBN statistics all-reduce over NCCL is the
    path's only data-path collective (plus the gradient all-reduce of DDP).  A step = one
    optimizer step over B frames per GPU.  `torch_syncbn` times the same module with
    torch.nn.SyncBatchNorm in place of the hrf_bn_* kernels."""
    import torch.nn as nn
    from hrfuser_b200 import HRFuserHRFormerBased, WORKLOADS, backbone_cfg, bn_train, ops, train
    from hrfuser_b200 import dist as hdist
    from hrfuser_b200.utils import randomize_parameters, synthetic_inputs
    rank, world, local = hdist.init_from_env()
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    variant, dataset, H, Wd = WORKLOADS[args.workload]
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)
    c = backbone_cfg(variant, dataset)
    c.pop('type')
    c['norm_cfg'] = dict(type='SyncBN', requires_grad=True)
    mod_ch = tuple(c.get('mod_in_channels', [3, 3]))

    def build(torch_bn):
        net = HRFuserHRFormerBased(**copy.deepcopy(c))
        randomize_parameters(net, 1)
        if torch_bn:
            for m in net.modules():
                if isinstance(m, bn_train.HrfSyncBatchNorm):
                    m.__class__ = nn.SyncBatchNorm
                elif isinstance(m, bn_train.HrfLayerNorm):
                    m.__class__ = nn.LayerNorm
        net = net.to(dev).train()
        if world > 1 and torch_bn:
            net = nn.parallel.DistributedDataParallel(net, device_ids=[local], find_unused_parameters=True)
        return net, torch.optim.SGD(net.parameters(), lr=1e-4, momentum=0.9)

    x, mods = synthetic_inputs(B, H, Wd, mod_ch, seed=10 + rank)
    x, mods = x.to(dev), [m.to(dev) for m in mods]
    loss_fn = lambda out: sum((o * o).mean() for o in out)

    def time_steps(step):
        for _ in range(Wm):
            step()
        torch.cuda.synchronize()
        hdist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = ops.launch_count()
        e0.record()
        for _ in range(K):
            step()
        e1.record()
        torch.cuda.synchronize()
        hdist.barrier()
        return hdist.max_over_ranks(e0.elapsed_time(e1), dev), ops.launch_count() - n0

    def eager(net, opt, manual_exchange):
        params = [p for p in net.parameters() if p.requires_grad]
        group = bn_train.sync_group(None)

        def step():
            opt.zero_grad(set_to_none=True)
            loss_fn(net(x, mods)).backward()
            if manual_exchange:
                train.exchange_gradients(params, group, world)
            opt.step()
        return time_steps(step)

    sampler = ClockSampler(local).start() if rank == 0 else None
    net, opt = build(False)
    ms_eager, launches_eager = eager(net, opt, True)       # the same step, launched op by op
    n0 = ops.launch_count()
    gstep = train.GraphedTrainStep(net, opt, x, mods, loss_fn)
    launches = (ops.launch_count() - n0) // 4              # 3 warm-up steps + the capture
    ms, _ = time_steps(gstep)
    loss = float(gstep.loss)
    clocks = sampler.stop() if sampler else None
    del net, opt, gstep
    torch.cuda.empty_cache()
    from hrfuser_b200 import modules as hmod
    hmod._WindowAttnBase.use_kernels = bn_train.HrfDepthwiseConv2d.use_kernels = False   # the torch arm: torch ops only
    net, opt = build(True)
    ms_t, _ = eager(net, opt, False)
    hmod._WindowAttnBase.use_kernels = bn_train.HrfDepthwiseConv2d.use_kernels = True
    if rank == 0:
        print(json.dumps({
            'metric': 'hrfuser_b_syncbn_train_frames_per_s', 'value': world * B * K / (ms / 1e3),
            'unit': 'frames/s', 'n_gpus': world, 'steps': K, 'warmup': Wm, 'ms_per_step': ms / K,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'gpu_launches': launches * K, 'hrf_kernel_launches_per_step': launches,
            'loss_last_step': loss,
            'config': {'workload': args.workload,
                       'mode': 'train (forward + backward + gradient exchange + SGD step) replayed as ONE CUDA graph '
                               '(hrfuser_b200.train.GraphedTrainStep)',
                       'frames_per_gpu_per_step': B, 'input': f'{H}x{Wd}', 'norm': 'SyncBN',
                       'parallelism': f'data parallel x{world}: SyncBN statistics all-reduce (one fp64 all-reduce of '
                                      '2C+1 values per BN and pass) + ONE flat-bucket gradient all-reduce, NCCL, '
                                      'all captured in the graph',
                       'weights': 'random-init (seeded)'},
            'eager': {'ms_per_step': ms_eager / K, 'frames_per_s': world * B * K / (ms_eager / 1e3),
                      'hrf_kernel_launches_per_step': launches_eager // K,
                      'what': 'the same step launched op by op from Python (no graph)'},
            'torch_syncbn': {'ms_per_step': ms_t / K, 'frames_per_s': world * B * K / (ms_t / 1e3),
                             'what': 'same module and step with torch.nn.SyncBatchNorm / nn.LayerNorm / torch attention core and depthwise convs (+ DistributedDataParallel '
                                     'when N > 1), eager'},
            'clocks': clocks,
        }))
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
    ap.add_argument('--no-eager-baseline', action='store_true')
    ap.add_argument('--train', action='store_true',
                    help='configs[3]: SyncBN training step (forward + backward + SGD), see run_train')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if not torch.cuda.is_available():
        sys.exit('bench.py: no CUDA device; the product path has no CPU fallback '
                 '(use --impl reference for the CPU port)')
    if args.train:
        return run_train(args)
    run_ours(args)


if __name__ == '__main__':
    main()
