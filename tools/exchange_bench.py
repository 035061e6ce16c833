"""GPU time of the exchange kernels (dwpw, pw, fuse_sum) per shape of the HRFuser-T stage-4
module, measured inside CUDA graphs of 20 calls.

    python tools/exchange_bench.py
"""
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hrfuser_b200 import ops  # noqa: E402
from hrfuser_b200.utils import randomize_parameters  # noqa: E402

B = 8
CH = [18, 36, 72, 144]
RES = [(96, 160), (48, 80), (24, 40), (12, 20)]


def graph_time(fn, n=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):
            fn(0)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(n):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5 / n * 1e3


def tok(i, c):
    H, W = RES[i]
    return [torch.randn(B, H, W, c, device='cuda').to(torch.bfloat16) for _ in range(4)]


for j in range(3):            # one down step from resolution j: dw3x3 s2 + bn + 1x1 + bn
    for cout in sorted({CH[j], CH[j + 1]}):
        dw, b1 = nn.Conv2d(CH[j], CH[j], 3, 2, 1, groups=CH[j], bias=False), nn.BatchNorm2d(CH[j])
        pw, b2 = nn.Conv2d(CH[j], cout, 1, bias=False), nn.BatchNorm2d(cout)
        randomize_parameters(nn.Sequential(dw, b1, pw, b2), 1)
        blob = ops.pack_dwpw(dw, b1, pw, b2).cuda()
        xs = tok(j, CH[j])
        t = graph_time(lambda i: ops.dw_down(xs[i % 4], blob, cout, True))
        print(f'dwpw  {RES[j][0]}x{RES[j][1]} C{CH[j]} -> C{cout}: {t:6.2f} us', flush=True)
for j in range(1, 4):         # 1x1 + bn at resolution j towards every finer branch
    for i in range(j):
        pw, bn = nn.Conv2d(CH[j], CH[i], 1, bias=False), nn.BatchNorm2d(CH[i])
        randomize_parameters(nn.Sequential(pw, bn), 1)
        blob = ops.pack_pw(pw, bn).cuda()
        xs = tok(j, CH[j])
        t = graph_time(lambda k: ops.pointwise(xs[k % 4], blob, CH[i]))
        print(f'pw    {RES[j][0]}x{RES[j][1]} C{CH[j]} -> C{CH[i]}: {t:6.2f} us', flush=True)
for i in range(4):            # fuse row i: x + (3 - i) up terms + i same-resolution terms
    xs = tok(i, CH[i])
    ups = [tok(j, CH[i])[0] for j in range(i + 1, 4)]
    sames = [tok(i, CH[i])[0] for _ in range(i)]
    for nchw in (False, True):
        t = graph_time(lambda k: ops.fuse_sum(xs[k % 4], ups, sames, relu=True, nchw_out=nchw))
        print(f'fuse  row {i} {RES[i][0]}x{RES[i][1]} C{CH[i]} ups {len(ups)} sames {len(sames)} nchw {int(nchw)}: {t:6.2f} us',
              flush=True)
