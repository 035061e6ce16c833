"""A few back-to-back calls of one op through the C-ABI (target for `ncu --set full`).
    python tools/op_once.py --kind lsa|mwca|mixffn [--C 18] [--M 2] [--n 4]"""
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
from microbench import GRIDS, WIDTHS, stub  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--kind', default='lsa')
ap.add_argument('--C', type=int, default=18)
ap.add_argument('--M', type=int, default=2)
ap.add_argument('--n', type=int, default=4)
ap.add_argument('--grid', default='nus')
a = ap.parse_args()
k = [w for w, _ in WIDTHS].index(a.C)
(H, W), (Cc, heads) = GRIDS[a.grid][k], WIDTHS[k]
e = stub()
xs = [torch.randn(8, H, W, Cc, device='cuda').to(torch.bfloat16) for _ in range(3)]
if a.kind == 'mixffn':
    blk, _ = make_block('lsa', Cc, heads)
    f = e._ffn(blk.norm2, blk.ffn)
    e._upload()
    fn = lambda i: ops.mixffn(xs[i], f['blob'].t, f['hidden'], f['eps'])
elif a.kind == 'lsa':
    blk, _ = make_block('lsa', Cc, heads)
    pk = e._hrformer_block(blk)
    e._upload()
    blobs = [s.t for s in pk['attn']]
    fn = lambda i: ops.window_attention(xs[i], None, blobs, heads)
else:
    blk, _ = make_block('mwca', Cc, heads, M=a.M)
    pk = e._fusion_block(blk)
    e._upload()
    blobs = [s.t for s in pk['attn']]
    fn = lambda i: ops.window_attention(xs[i], [xs[(i + 1 + m) % 3] for m in range(a.M)], blobs, heads)
for i in range(a.n):
    fn(i % 3)
torch.cuda.synchronize()
print('done')
