"""HBM throughput of the train-mode BatchNorm kernels (hrf_bn_stats / hrf_bn_normalize /
hrf_bn_bwd_stats / hrf_bn_bwd_dx) at the training shapes of the backbone, against MEASURED_PEAKS.json.

Each shape rotates over enough input sets to exceed the 126 MB L2; times are CUDA events
around a CUDA graph of the rotating calls (no launch gaps).

    python tools/bn_bench.py [--json out.jsonl]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hrfuser_b200 import ops  # noqa: E402

SHAPES = [  # (label, B, C, H, W)
    ('stem 64ch 192x320 B8', 8, 64, 192, 320),
    ('layer1 256ch 96x160 B8', 8, 256, 96, 160),
    ('ffn hidden 72ch 96x160 B8', 8, 72, 96, 160),
    ('branch0 18ch 96x160 B8', 8, 18, 96, 160),
    ('ffn hidden 312ch 96x160 B2 (HRFuser-B)', 2, 312, 96, 160),
    ('branch3 144ch 12x20 B8', 8, 144, 12, 20),
]


def graph_time_us(fn, n_sets, reps=4):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(n_sets):
            fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(n_sets):
                fn(i)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / n_sets * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--json')
    ap.add_argument('--dtype', default='fp32', choices=['fp32', 'bf16'])
    a = ap.parse_args()
    dt = torch.float32 if a.dtype == 'fp32' else torch.bfloat16
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    rows = []
    for label, B, C, H, W in SHAPES:
        nbytes = B * C * H * W * (4 if dt == torch.float32 else 2)
        n_sets = max(2, int(400e6 // nbytes) + 1)           # > 3 x L2 of distinct inputs
        n_sets = min(n_sets, 64)
        xs = [torch.randn(B, C, H, W, device='cuda').to(dt) for _ in range(n_sets)]
        dys = [torch.randn(B, C, H, W, device='cuda').to(dt) for _ in range(n_sets)]
        mean = torch.zeros(C, device='cuda')
        invstd = torch.ones(C, device='cuda')
        w, bias = torch.ones(C, device='cuda'), torch.zeros(C, device='cuda')
        rm, rv = torch.zeros(C, device='cuda'), torch.ones(C, device='cuda')
        stats = ops.bn_stats(xs[0])
        sums = ops.bn_bwd_stats(xs[0], dys[0], mean, invstd)
        cases = {
            'bn_stats': (lambda i: ops.bn_stats(xs[i]), 1),
            'bn_normalize': (lambda i: ops.bn_normalize(xs[i], stats, w, bias, 1e-5, 0.1, rm, rv), 2),
            'bn_bwd_stats': (lambda i: ops.bn_bwd_stats(xs[i], dys[i], mean, invstd), 2),
            'bn_bwd_dx': (lambda i: ops.bn_bwd_dx(xs[i], dys[i], sums, stats[2 * C:], w, mean,
                                                  invstd), 3),
        }
        for name, (fn, passes) in cases.items():
            us = graph_time_us(fn, n_sets)
            gbs = passes * nbytes / us / 1e3
            row = dict(kernel=name, shape=label, dtype=a.dtype, us=round(us, 2),
                       algorithmic_bytes=passes * nbytes, gbs=round(gbs, 1),
                       frac_of_hbm_peak=round(gbs / peak, 3), peak_gbs=peak,
                       note='stats = reduce + finalize launches')
            rows.append(row)
            print(f'{name:14s} {label:42s} {us:9.2f} us  {gbs:8.1f} GB/s  {gbs / peak:6.1%} of {peak:.0f}',
                  flush=True)
        del xs, dys
        torch.cuda.empty_cache()
    if a.json:
        with open(a.json, 'w') as f:
            for r in rows:
                f.write(json.dumps(r) + '\n')


if __name__ == '__main__':
    main()
