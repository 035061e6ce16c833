"""Graph-timed LSA and MixFFN kernels (B=8, one HRNet branch) for the CTAs-per-SM sweep:

    for n in 1 2 3 4; do HRF_ATTN_CTAS_PER_SM=$n HRF_FFN_CTAS_PER_SM=$n python tools/occupancy_sweep.py; done
"""
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


def graph_time(fn, n=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
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


out = []
for (H, W), (C, heads) in zip(GRIDS['nus'], WIDTHS):
    e = stub()
    blk, _ = make_block('lsa', C, heads)
    f = e._ffn(blk.norm2, blk.ffn)
    pk = e._hrformer_block(blk)
    e._upload()
    blobs = [s.t for s in pk['attn']]
    xs = [torch.randn(8, H, W, C, device='cuda').to(torch.bfloat16) for _ in range(4)]
    t_l = graph_time(lambda i: ops.window_attention(xs[i % 4], None, blobs, heads))
    t_f = graph_time(lambda i: ops.mixffn(xs[i % 4], f['blob'].t, f['hidden'], f['eps']))
    out.append(f'C={C}: lsa {t_l:5.1f} ffn {t_f:5.1f}')
print(f"attn/ffn CTAs per SM = {os.environ.get('HRF_ATTN_CTAS_PER_SM', 'default')}/"
      f"{os.environ.get('HRF_FFN_CTAS_PER_SM', 'default')} (us):  " + '   '.join(out))
