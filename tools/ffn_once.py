"""One MixFFN call through the C-ABI, second-generation kernel against the first (debug aid).
    python tools/ffn_once.py [--C 18] [--H 96] [--W 160] [--B 2]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import make_block, tokens  # noqa: E402
from hrfuser_b200 import ops  # noqa: E402
from hrfuser_b200.engine import BackboneEngine  # noqa: E402
from oracle import hrfuser_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--C', type=int, default=18)
    ap.add_argument('--heads', type=int, default=1)
    ap.add_argument('--H', type=int, default=96)
    ap.add_argument('--W', type=int, default=160)
    ap.add_argument('--B', type=int, default=2)
    a = ap.parse_args()
    blk, sd = make_block('lsa', a.C, a.heads, seed=7)
    e = BackboneEngine.__new__(BackboneEngine)
    e._host_blobs, e._blob_slots = [], []
    e.device, e.dtype = torch.device('cuda'), torch.float32
    f = e._ffn(blk.norm2, blk.ffn)
    e._upload()
    x = tokens(a.B, a.H, a.W, a.C, seed=5).to(torch.bfloat16)
    xc = x.cuda()
    with torch.no_grad():
        t = x.float().view(a.B, a.H * a.W, a.C)
        ref = (t + O.cross_ffn(O.layer_norm(t, sd, 'blk.norm2'), sd, 'blk.ffn', a.H, a.W)).view_as(x)
    res = {}
    for v in ('0', '1'):
        os.environ['HRF_FFN_V2'] = v
        y = ops.mixffn(xc, f['blob'].t, f['hidden'], f['eps'])
        torch.cuda.synchronize()
        res[v] = y.float().cpu()
        err = float((res[v] - ref).norm() / ref.norm())
        bad = (~torch.isclose(res[v], ref, rtol=2e-2, atol=2e-2 * float(ref.pow(2).mean().sqrt()))).float().mean()
        print(f'v2={v}: norm-wise error vs oracle {err:.3e}, elements out of tolerance {float(bad):.3%}', flush=True)
    d = (res['1'] - res['0']).abs()
    print('v2 vs v1: max abs diff', float(d.max()), 'at', [int(i) for i in torch.unravel_index(d.argmax(), d.shape)])


if __name__ == '__main__':
    main()
