"""hrf_convgemm_fwd against cuDNN (the fused cudnn_convolution_relu / _add_relu calls the engine
used before) on the stem / Bottleneck / transition / HRFPN layer shapes, B = 8, bf16.

    python tools/convgemm_bench.py [--batch 8] [--iters 30]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import peaks  # noqa: E402
from hrfuser_b200 import ops  # noqa: E402
from hrfuser_b200.engine import _Conv  # noqa: E402
from hrfuser_b200.utils import randomize_parameters  # noqa: E402

LAYERS = [  # name, Cin, Cout, k, stride, relu, residual, (H, W) of the INPUT
    ('stem conv2', 64, 64, 3, 2, True, False, (192, 320)),
    ('bottleneck conv1 (64)', 64, 64, 1, 1, True, False, (96, 160)),
    ('bottleneck conv1 (256)', 256, 64, 1, 1, True, False, (96, 160)),
    ('bottleneck conv2', 64, 64, 3, 1, True, False, (96, 160)),
    ('bottleneck conv3 + add', 64, 256, 1, 1, True, True, (96, 160)),
    ('downsample', 64, 256, 1, 1, False, False, (96, 160)),
    ('transition 256->18', 256, 18, 3, 1, False, False, (96, 160)),
    ('transition 256->36 s2', 256, 36, 3, 2, True, False, (96, 160)),
    ('hrfpn 256->256', 256, 256, 3, 1, False, False, (96, 160)),
    ('hrfpn 256->256 /2', 256, 256, 3, 1, False, False, (48, 80)),
]


def timed(fn, iters):
    """graph-timed (tools/microbench.py): the ctypes launch path (~20 us) stays out of the number"""
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    from microbench import timed as graph_timed
    return graph_timed(lambda i: fn(), 1, iters)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--iters', type=int, default=30)
    ap.add_argument('--group', type=int, default=3, help='streams per grouped launch (1 + M)')
    a = ap.parse_args()
    pk = peaks()
    B = a.batch
    for name, cin, cout, k, stride, relu, resid, (H, W) in LAYERS:
        conv = nn.Conv2d(cin, cout, k, stride, k // 2, bias=False)
        bn = nn.BatchNorm2d(cout)
        randomize_parameters(nn.Sequential(conv, bn), 1)
        bn.eval()
        blob = ops.pack_convgemm(conv, bn, bn.eps).cuda()
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        n_sets = 6            # rotate buffers: > 126 MB L2 for the 256-channel maps
        xs = [torch.randn(B, H, W, cin, device='cuda').to(torch.bfloat16) for _ in range(n_sets)]
        rs = [torch.randn(B, Ho, Wo, cout, device='cuda').to(torch.bfloat16) for _ in range(n_sets)] if resid else None
        i = [0]

        G = a.group

        def ours():
            j = i[0] = (i[0] + 1) % n_sets
            if G == 1:
                return ops.conv_gemm(xs[j], blob, cout, k, stride, relu, rs[j] if resid else None)
            sel = [(j + q) % n_sets for q in range(G)]
            return ops.conv_gemm_grouped([xs[q] for q in sel], [blob] * G, cout, k, stride, [relu] * G,
                                         [rs[q] for q in sel] if resid else None)
        cv = _Conv(conv.cuda(), bn.cuda(), relu, torch.bfloat16)

        def cudnn():
            for _ in range(G):
                j = i[0] = (i[0] + 1) % n_sets
                cv(xs[j].permute(0, 3, 1, 2), rs[j].permute(0, 3, 1, 2) if resid else None)
        t_o, t_c = timed(ours, a.iters) / G, timed(cudnn, a.iters) / G
        byts = (B * H * W * cin + B * Ho * Wo * cout * (2 if resid else 1)) * 2
        flops = 2 * B * Ho * Wo * k * k * cin * cout
        print(json.dumps(dict(layer=name, group=G, B=B, H=H, W=W, cin=cin, cout=cout, k=k, stride=stride,
                              convgemm_ms=round(t_o, 5), cudnn_ms=round(t_c, 5),
                              GBps=round(byts / t_o / 1e6, 1), frac_hbm=round(byts / t_o / 1e6 / pk['hbm_gbs'], 4),
                              TFLOPs=round(flops / t_o / 1e9, 2),
                              frac_tensor=round(flops / t_o / 1e9 / pk['bf16_tflops'], 4))), flush=True)


if __name__ == '__main__':
    main()
