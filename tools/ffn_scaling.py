"""MixFFN / LSA kernel time vs problem size (fixed overhead vs per-tile cost).

    python tools/ffn_scaling.py [--C 18]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from helpers import make_block  # noqa: E402
from hrfuser_b200 import ops  # noqa: E402
from microbench import stub, timed  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--C', type=int, default=18)
ap.add_argument('--heads', type=int, default=1)
a = ap.parse_args()
e = stub()
blk, _ = make_block('lsa', a.C, a.heads)
f = e._ffn(blk.norm2, blk.ffn)
pk = e._hrformer_block(blk)
e._upload()
blobs = [s.t for s in pk['attn']]
for (B, H, W) in [(1, 6, 14), (1, 42, 56), (1, 96, 160), (2, 96, 160), (4, 96, 160), (8, 96, 160), (16, 96, 160)]:
    xs = [torch.randn(B, H, W, a.C, device='cuda').to(torch.bfloat16) for _ in range(4)]
    t_ffn = timed(lambda i: ops.mixffn(xs[i], f['blob'].t, f['hidden'], f['eps']), 4, 50)
    t_lsa = timed(lambda i: ops.window_attention(xs[i], None, blobs, a.heads), 4, 50)
    print(f'B={B} {H}x{W}: tokens {B * H * W:7d}  mixffn {t_ffn * 1e3:7.2f} us   lsa {t_lsa * 1e3:7.2f} us', flush=True)

# the same back-to-back chain inside a CUDA graph (what the engine replays)
B, H, W = 8, 96, 160
xs = [torch.randn(B, H, W, a.C, device='cuda').to(torch.bfloat16) for _ in range(2)]
outs = [torch.empty_like(xs[0]) for _ in range(2)]
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        ops.mixffn(xs[0], f['blob'].t, f['hidden'], f['eps'], out=outs[0])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for i in range(20):       # alternating lsa -> ffn chain like one HRNet branch
            ops.window_attention(xs[i % 2], None, blobs, a.heads)
            ops.mixffn(xs[i % 2], f['blob'].t, f['hidden'], f['eps'], out=outs[i % 2])
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f'graph of 20 x (lsa + mixffn) at B=8: {e0.elapsed_time(e1) / 10 / 20 * 1e3:.2f} us per pair')
