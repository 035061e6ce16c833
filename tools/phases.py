"""Per-phase cycle breakdown of the tcgen05 kernels (instrumented build, -DHRF_KERNEL_PROFILE).

    bash tools/gpu_phases.sh            # builds the instrumented library here, runs this on the GPU
    HRF_LIB=.../libhrfuser_b200_prof.so python tools/phases.py --kind ffn --C 18

Thread 0 of every CTA accumulates clock64 deltas per phase (common.cuh); this prints the mean
cycles per tile of each phase over all CTAs, and the per-CTA setup time.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from helpers import make_block  # noqa: E402
from hrfuser_b200 import _lib, ops  # noqa: E402
from microbench import GRIDS, WIDTHS, stub  # noqa: E402

NAMES = {
    'ffn': ['LN prologue', 'sync', 'fc1 issue', 'fc1 wait', 'epilogue 1', 'sync', 'dw conv', 'sync',
            'fc2 issue', 'fc2 wait', 'epilogue 2', '-', '-', '-'],
    'ffn_tcd': ['LN prologue', 'sync', 'fc1 issue', 'fc1 wait', 'epilogue 1', 'sync', 'conv issue',
                'conv wait', 'epilogue dw', 'sync', 'fc2 issue', 'fc2 wait', 'epilogue 2', '-'],
    'attn': ['LN prologue', 'sync', 'qkv issue', 'qkv wait', 'qkv/PV epilogue', 'sync+S issue', 'S wait',
             'softmax', 'sync+PV issue', 'PV wait', 'PV epi (last)', 'sync+out issue', 'out wait',
             'out epilogue'],
}
ap = argparse.ArgumentParser()
ap.add_argument('--kind', default='ffn', choices=['ffn', 'lsa', 'mwca'])
ap.add_argument('--C', type=int, default=18)
ap.add_argument('--batch', type=int, default=8)
a = ap.parse_args()
lib = _lib.load()
lib.hrf_debug_prof.argtypes = [C.c_void_p, C.c_int, C.c_int]
k = [w for w, _ in WIDTHS].index(a.C)
(H, W), (Cc, heads) = GRIDS['nus'][k], WIDTHS[k]
e = stub()
x = torch.randn(a.batch, H, W, Cc, device='cuda').to(torch.bfloat16)
if a.kind == 'ffn':
    blk, _ = make_block('lsa', Cc, heads)
    f = e._ffn(blk.norm2, blk.ffn)
    e._upload()
    fn = lambda: ops.mixffn(x, f['blob'].t, f['hidden'], f['eps'])
    names = NAMES['ffn_tcd' if os.environ.get('HRF_FFN_DW_TC') == '1' and Cc == 18 else 'ffn']
elif a.kind == 'lsa':
    blk, _ = make_block('lsa', Cc, heads)
    pk = e._hrformer_block(blk)
    e._upload()
    blobs = [s.t for s in pk['attn']]
    fn = lambda: ops.window_attention(x, None, blobs, heads)
    names = NAMES['attn']
else:
    blk, _ = make_block('mwca', Cc, heads, M=2)
    pk = e._fusion_block(blk)
    e._upload()
    blobs = [s.t for s in pk['attn']]
    zs = [torch.randn_like(x) for _ in range(2)]
    fn = lambda: ops.window_attention(x, zs, blobs, heads)
    names = NAMES['attn']
for it in range(3):
    fn()
torch.cuda.synchronize()
lib.hrf_debug_prof(None, 0, 1)
n_it = 10
for it in range(n_it):
    fn()
torch.cuda.synchronize()
buf = np.zeros(2048 * 16, dtype=np.uint64)
lib.hrf_debug_prof(buf.ctypes.data, buf.size, 0)
t = buf.reshape(2048, 16).astype(np.float64)
t = t[t[:, 15] > 0]
tiles = t[:, 15].sum()
print(f'{a.kind} C={Cc} {H}x{W} B={a.batch}: {len(t)} CTAs, {tiles / n_it:.0f} tiles per launch, '
      f'{t[:, 15].mean() / n_it:.2f} tiles per CTA')
tot = 0.0
for j in range(14):
    v = t[:, j].sum() / tiles
    tot += v
    if v > 0:
        print(f'  {names[j]:16s} {v:9.0f} cycles / tile')
print(f'  {"total":16s} {tot:9.0f} cycles / tile;  setup {t[:, 14].mean() / n_it:.0f} cycles / CTA')
