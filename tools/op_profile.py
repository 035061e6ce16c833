"""Kernel durations (torch.profiler / CUPTI) of one op call per HRFuser-B width (debug aid).
    python tools/op_profile.py [--variant b]"""
import argparse
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from helpers import make_block  # noqa: E402
from hrfuser_b200 import ops  # noqa: E402
from microbench import GRIDS, WIDTHS, WIDTHS_B, stub  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--variant', default='b')
a = ap.parse_args()
for (H, W), (C, heads) in zip(GRIDS['nus'], WIDTHS if a.variant == 't' else WIDTHS_B):
    e = stub()
    blk, _ = make_block('lsa', C, heads)
    pk = e._hrformer_block(blk)
    e._upload()
    x = torch.randn(8, H, W, C, device='cuda').to(torch.bfloat16)
    blobs = [s.t for s in pk['attn']]
    f = pk['ffn']
    for kind, fn in (('lsa', lambda: ops.window_attention(x, None, blobs, heads)),
                     ('mixffn', lambda: ops.mixffn(x, f['blob'].t, f['hidden'], f['eps']))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        rows = [(ev.name.replace('void hrf::', '')[:60], ev.time_range.end - ev.time_range.start)
                for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA]
        print(f'{kind} C={C} {H}x{W}: ' + ', '.join(f'{n.split("(")[0]} {d:.0f}us' for n, d in rows))
