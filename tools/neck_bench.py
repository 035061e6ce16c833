"""HRFPN front half (upsample x3 + concat + 1x1 reduction, hrfpn.py:79-86) on the B200:
hrfuser_b200.neck.HRFPN.reduce (4 x hrf_pw_fwd at branch resolution + hrf_fuse_sum_fwd) against
the reference's formulation in torch eager on the same GPU (its own execution model).

    python tools/neck_bench.py [--json out.jsonl]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hrfuser_b200.neck import HRFPN  # noqa: E402


def time_ms(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--json')
    a = ap.parse_args()
    rows = []
    for label, chans, grid, B in (('HRFuser-T nus 8 frames', [18, 36, 72, 144], (96, 160), 8),
                                  ('HRFuser-T stf 8 frames', [18, 36, 72, 144], (96, 312), 8),
                                  ('HRFuser-B nus 8 frames', [78, 156, 312, 624], (96, 160), 8)):
        torch.manual_seed(0)
        xs = [torch.randn(B, c, grid[0] >> i, grid[1] >> i, device='cuda') for i, c in enumerate(chans)]
        row = dict(shape=label)
        for mode in ('fp32', 'bf16'):
            net = HRFPN(in_channels=chans, out_channels=256, precision=mode).eval().cuda()
            with torch.no_grad():
                row[f'ours_{mode}_ms'] = round(time_ms(lambda: net.reduce(xs)), 4)
        conv = net.reduction_conv

        def eager():
            outs = [xs[0]] + [F.interpolate(xs[i], scale_factor=2 ** i, mode='bilinear') for i in range(1, 4)]
            return conv(torch.cat(outs, dim=1))

        with torch.no_grad():
            row['torch_eager_fp32_ms'] = round(time_ms(eager), 4)
            with torch.autocast('cuda', dtype=torch.bfloat16):
                row['torch_eager_bf16_autocast_ms'] = round(time_ms(eager), 4)
        # the whole neck (front half + pooling pyramid + five 3x3 convs -> 5 fp32 NCHW maps)
        for mode in ('fp32', 'bf16'):
            net = HRFPN(in_channels=chans, out_channels=256, precision=mode).eval().cuda()
            with torch.no_grad():
                row[f'whole_ours_{mode}_ms'] = round(time_ms(lambda: net(xs)), 4)
        with torch.no_grad():
            row['whole_torch_eager_fp32_ms'] = round(time_ms(lambda: net._forward_autograd(xs)), 4)
            with torch.autocast('cuda', dtype=torch.bfloat16):
                row['whole_torch_eager_bf16_autocast_ms'] = round(time_ms(lambda: net._forward_autograd(xs)), 4)
        n0 = B * grid[0] * grid[1]
        row['concat_tensor_MB_avoided'] = round(n0 * sum(chans) * 4 / 1e6, 1)
        row['gflop_ours'] = round(2 * 256 * sum((n0 >> (2 * i)) * c for i, c in enumerate(chans)) / 1e9, 2)
        row['gflop_reference'] = round(2 * 256 * n0 * sum(chans) / 1e9, 2)
        rows.append(row)
        print(json.dumps(row), flush=True)
    if a.json:
        with open(a.json, 'w') as f:
            for r in rows:
                f.write(json.dumps(r) + '\n')


if __name__ == '__main__':
    main()
