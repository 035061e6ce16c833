"""HBM throughput of the input prologue (hrf_input_prologue_fwd) at the frame sizes of the
BASELINE.json configs, against MEASURED_PEAKS.json, with the reference's CPU pipeline
(oracle/input_oracle.py = its numpy / OpenCV arithmetic) timed beside it on one frame batch.

Each shape rotates over enough source / destination sets to exceed the 126 MB L2; times are
CUDA events around a CUDA graph of the rotating calls.  Algorithmic bytes per launch =
B*H*W*C*sizeof(src) + B*C*Hp*Wp*4.

    python tools/input_bench.py [--json out.jsonl] [--once]   (--once: one call per shape, for ncu)
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hrfuser_b200 import ops  # noqa: E402
from tools.bn_bench import graph_time_us  # noqa: E402

IMG = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)
LIDAR = dict(mean=[0.23277158, 0.31501067, -0.00012928071],
             std=[2.5538357826888602, 3.7345728854535643, 0.2815488539921788], to_rgb=False)
GATED = dict(mean=[181.74427536], std=[185.49071888], to_rgb=False)
RADAR2 = dict(mean=[3.4423912, 0.021001821], std=[19.330362993097626, 0.7612592077132296], to_rgb=False)
SHAPES = [  # label, (B,H,W,C), dtype, cfg
    ('nus camera u8 8x360x640x3', (8, 360, 640, 3), torch.uint8, IMG),
    ('nus lidar f32 8x360x640x3', (8, 360, 640, 3), torch.float32, LIDAR),
    ('stf camera u8 8x384x1248x3', (8, 384, 1248, 3), torch.uint8, IMG),
    ('stf radar f32 8x384x1248x2', (8, 384, 1248, 2), torch.float32, RADAR2),
    ('stf gated u8 8x384x1248x1', (8, 384, 1248, 1), torch.uint8, GATED),
    ('camera u8 64x360x640x3 (8 steps batched)', (64, 360, 640, 3), torch.uint8, IMG),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--json')
    ap.add_argument('--once', action='store_true')
    a = ap.parse_args()
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
    rows = []
    for label, shape, dt, cfg in SHAPES:
        B, H, W, Cc = shape
        Hp, Wp = -(-H // 32) * 32, -(-W // 32) * 32
        nbytes = B * H * W * Cc * (1 if dt == torch.uint8 else 4) + B * Cc * Hp * Wp * 4
        n_sets = 1 if a.once else max(2, -(-3 * 126_000_000 // nbytes))
        if dt == torch.uint8:
            srcs = [torch.randint(0, 256, shape, dtype=dt, device='cuda') for _ in range(n_sets)]
        else:
            srcs = [torch.randn(shape, device='cuda') * 20 for _ in range(n_sets)]
        dsts = [torch.empty(B, Cc, Hp, Wp, device='cuda') for _ in range(n_sets)]

        def fn(i):
            ops.input_prologue(srcs[i], cfg['mean'], cfg['std'], to_rgb=cfg['to_rgb'], out=dsts[i])

        if a.once:
            fn(0)
            torch.cuda.synchronize()
            continue
        us = graph_time_us(fn, n_sets)
        row = dict(shape=label, us=round(us, 2), alg_bytes=nbytes, gbs=round(nbytes / us * 1e-3, 1),
                   frac_of_hbm_peak=round(nbytes / us * 1e-3 / peak, 3), peak_gbs=peak, sets=n_sets)
        rows.append(row)
        print(json.dumps(row), flush=True)
        del srcs, dsts
    if not a.once:
        # the reference's CPU arithmetic on the first shape (one batch), host cores as numpy uses them
        from oracle import input_oracle
        frames = np.random.default_rng(0).integers(0, 256, SHAPES[0][1], dtype=np.uint8)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            input_oracle.input_prologue(frames, IMG['mean'], IMG['std'], to_rgb=True)
        cpu_ms = (time.perf_counter() - t0) / reps * 1e3
        row = dict(shape=SHAPES[0][0], cpu_port_ms=round(cpu_ms, 2), gpu_us=rows[0]['us'],
                   note='numpy restatement of Normalize+Pad+FormatBundle, 1 thread')
        rows.append(row)
        print(json.dumps(row))
    if a.json:
        with open(a.json, 'w') as f:
            for r in rows:
                f.write(json.dumps(r) + '\n')


if __name__ == '__main__':
    main()
