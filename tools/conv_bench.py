"""GPU time of the transition conv kernel (hrf_conv3x3_fwd) per shape, measured inside a CUDA
graph of 20 calls (no host launch overhead in the number).

    python tools/conv_bench.py
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
SHAPES = [(18, 18, 96, 160), (18, 36, 96, 160), (18, 18, 48, 80), (18, 72, 48, 80), (18, 18, 24, 40),
          (18, 144, 24, 40), (36, 72, 48, 80), (72, 144, 24, 40)]
for cin, cout, H, W in SHAPES:
    conv, bn = nn.Conv2d(cin, cout, 3, 2, 1, bias=False), nn.BatchNorm2d(cout)
    randomize_parameters(nn.Sequential(conv, bn), 1)
    blob = ops.pack_conv3x3(conv, bn, bn.eps).cuda()
    xs = [torch.randn(B, H, W, cin, device='cuda').to(torch.bfloat16) for _ in range(4)]
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):
            ops.conv3x3(xs[0], blob, cout, 2, True)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(20):
                ops.conv3x3(xs[i % 4], blob, cout, 2, True)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 200 * 1e3
    no = B * ((H + 1) // 2) * ((W + 1) // 2)
    print(f'{cin:3d} -> {cout:3d}  in {H}x{W}  out tokens {no:6d}: {us:7.2f} us  '
          f'({2 * no * 9 * cin * cout / us / 1e6:6.2f} TFLOP/s)', flush=True)
