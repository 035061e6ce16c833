"""Kernel-time breakdown of one HRFuser-B training step (torch.profiler / CUPTI), by kernel family.
    python tools/train_profile.py [--batch 2] [--top 40]"""
import argparse
import collections
import copy
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hrfuser_b200 import HRFuserHRFormerBased, WORKLOADS, backbone_cfg  # noqa: E402
from hrfuser_b200.utils import randomize_parameters, synthetic_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='hrfuser_b_nus_r640')
ap.add_argument('--batch', type=int, default=2)
ap.add_argument('--top', type=int, default=40)
a = ap.parse_args()
dev = torch.device('cuda', 0)
variant, dataset, H, W = WORKLOADS[a.workload]
c = backbone_cfg(variant, dataset)
c.pop('type')
c['norm_cfg'] = dict(type='SyncBN', requires_grad=True)
net = HRFuserHRFormerBased(**copy.deepcopy(c))
randomize_parameters(net, 1)
net = net.to(dev).train()
opt = torch.optim.SGD(net.parameters(), lr=1e-4, momentum=0.9)
x, mods = synthetic_inputs(a.batch, H, W, tuple(c.get('mod_in_channels', [3, 3])), seed=1)
x, mods = x.to(dev), [m.to(dev) for m in mods]


def step():
    opt.zero_grad(set_to_none=True)
    out = net(x, mods)
    sum((o * o).mean() for o in out).backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
tot = collections.Counter()
cnt = collections.Counter()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = e.name.replace('(anonymous namespace)::', '').replace('void ', '').split('<')[0].split('(')[0][:80]
        d = e.time_range.end - e.time_range.start
        tot[name] += d
        cnt[name] += 1
s = sum(tot.values())
print(f'{sum(cnt.values())} kernels, sum of durations {s / 1e3:.1f} ms')
for n, t in tot.most_common(a.top):
    print(f'{t / 1e3:9.2f} ms {100 * t / s:5.1f}%  x{cnt[n]:<5d} {n}')

# ---- the same step by the op that launched the kernels (autograd nodes / aten ops, self device time)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof2:
    step()
    torch.cuda.synchronize()
rows = []
for k in prof2.key_averages():
    t = getattr(k, 'self_device_time_total', None)
    if t is None:
        t = getattr(k, 'self_cuda_time_total', 0)
    if t > 0:
        rows.append((t, k.count, k.key))
rows.sort(reverse=True)
s2 = sum(r[0] for r in rows)
print(f'\nby op (self device time), total {s2 / 1e3:.1f} ms')
for t, n, key in rows[:a.top]:
    print(f'{t / 1e3:9.2f} ms {100 * t / s2:5.1f}%  x{n:<5d} {key[:90]}')
