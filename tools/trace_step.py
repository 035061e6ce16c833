"""Kernel timeline of graph replays of the backbone step (torch.profiler / CUPTI):

    python tools/trace_step.py [--out gpurun_out/trace.json]

Writes one record per kernel of the LAST profiled replay: name, start (us), duration (us),
stream.  tools/trace_report.py turns it into concurrency / critical-path statistics.
"""
import argparse
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_net  # noqa: E402
from hrfuser_b200.engine import GraphedForward  # noqa: E402
from hrfuser_b200.utils import synthetic_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='hrfuser_t_nus_r640')
ap.add_argument('--batch', type=int, default=8)
ap.add_argument('--precision', default='bf16')
ap.add_argument('--out', default='gpurun_out/trace.json')
a = ap.parse_args()
dev = torch.device('cuda', 0)
cfg, net, (H, W), mod_ch = build_net(a.workload, a.precision, dev)
x, mods = synthetic_inputs(a.batch, H, W, mod_ch, seed=0, device=dev)
with torch.no_grad():
    eng = net.engine()
    g = GraphedForward(eng, x, mods)
    for _ in range(5):
        g()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            g()
            torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
recs = [{'name': e.name, 'start': e.time_range.start, 'dur': e.time_range.end - e.time_range.start,
         'stream': getattr(e, 'device_resource_id', -1)} for e in ev]
recs.sort(key=lambda r: r['start'])
os.makedirs(os.path.dirname(a.out) or '.', exist_ok=True)
json.dump(recs, open(a.out, 'w'))
print(len(recs), 'device events')
