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
    'attn_v3': ['LN + prefetch', 'sync', 'proj issue', 'proj wait', 'proj epilogue', 'sync+S issue', 'S wait',
                'softmax', 'sync+PV issue', 'PV wait', 'out epilogue', '-', '-', '-'],
    'ffn': ['LN prologue', 'sync', 'fc1 issue', 'fc1 wait', 'epilogue 1', 'sync', 'dw conv', 'sync',
            'fc2 issue', 'fc2 wait', 'epilogue 2', '-', '-', '-'],
    'ffn_v2': ['TMA wait + LN', 'sync', 'fc1 issue', 'fc1 wait', 'epilogue 1', 'sync', 'dw conv', 'sync',
               'fc2 issue', 'fc2 wait', 'epilogue 2', 'sync + store', '-', '-'],
    'attn': ['LN prologue', 'sync', 'qkv issue', 'qkv wait', 'qkv/PV epilogue', 'sync+S issue', 'S wait',
             'softmax', 'sync+PV issue', 'PV wait', 'PV epi (last)', 'sync+out issue', 'out wait',
             'out epilogue'],
}
ap = argparse.ArgumentParser()
ap.add_argument('--kind', default='ffn', choices=['ffn', 'lsa', 'mwca', 'stem'])
ap.add_argument('--C', type=int, default=18)
ap.add_argument('--batch', type=int, default=8)
a = ap.parse_args()
lib = _lib.load()
lib.hrf_debug_prof.argtypes = [C.c_void_p, C.c_int, C.c_int]
if a.kind == 'stem':
    a.C = 18
k = [w for w, _ in WIDTHS].index(a.C)
(H, W), (Cc, heads) = GRIDS['nus'][k], WIDTHS[k]
e = stub()
x = torch.randn(a.batch, H, W, Cc, device='cuda').to(torch.bfloat16)
if a.kind == 'stem':
    import torch.nn as nn
    from hrfuser_b200.utils import randomize_parameters
    conv, bn = nn.Conv2d(3, 64, 3, 2, 1, bias=False), nn.BatchNorm2d(64)
    randomize_parameters(nn.Sequential(conv, bn), 3)
    sblob = ops.pack_stem(conv, bn, bn.eps).cuda()
    img = torch.randn(a.batch, 3, 384, 640, device='cuda')
    fn = lambda: ops.stem_conv(img, sblob, 64)
    names = ['patch -> tile', 'sync', 'MMA issue', 'next loads', 'MMA wait', 'epilogue', 'sync'] + ['-'] * 7
elif a.kind == 'ffn':
    blk, _ = make_block('lsa', Cc, heads)
    f = e._ffn(blk.norm2, blk.ffn)
    e._upload()
    fn = lambda: ops.mixffn(x, f['blob'].t, f['hidden'], f['eps'])
    names = NAMES['ffn_v2' if os.environ.get('HRF_FFN_V2', '1') != '0' and Cc == 18 else 'ffn']
elif a.kind == 'lsa':
    blk, _ = make_block('lsa', Cc, heads)
    pk = e._hrformer_block(blk)
    e._upload()
    blobs = [s.t for s in pk['attn']]
    fn = lambda: ops.window_attention(x, None, blobs, heads)
    names = NAMES['attn_v3' if Cc == 18 and os.environ.get('HRF_ATTN_V3', '1') != '0' else 'attn']
else:
    blk, _ = make_block('mwca', Cc, heads, M=2)
    pk = e._fusion_block(blk)
    e._upload()
    blobs = [s.t for s in pk['attn']]
    zs = [torch.randn_like(x) for _ in range(2)]
    fn = lambda: ops.window_attention(x, zs, blobs, heads)
    names = NAMES['attn_v3' if Cc == 18 and os.environ.get('HRF_ATTN_V3', '1') != '0' else 'attn']
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

# residency: SM id and globaltimer lifetime of every CTA of the LAST launch
lib.hrf_debug_prof_cta.argtypes = [C.c_void_p, C.c_int]
cb = np.zeros(2048 * 4, dtype=np.uint64)
lib.hrf_debug_prof_cta(cb.ctypes.data, cb.size)
cta = cb.reshape(2048, 4)[:len(t)].astype(np.int64)
t_begin = cta[:, 1].min()
span = (cta[:, 2].max() - t_begin) / 1e3
per_sm = {}
for smid, a0, a1, _ in cta:
    per_sm.setdefault(int(smid), []).append((int(a0 - t_begin), int(a1 - t_begin)))
peak = []
for smid, iv in per_sm.items():
    ev = sorted([(a0, 1) for a0, _ in iv] + [(a1, -1) for _, a1 in iv])
    cur = mx = 0
    for _, d in ev:
        cur += d
        mx = max(mx, cur)
    peak.append(mx)
life = (cta[:, 2] - cta[:, 1]) / 1e3
starts = np.sort(cta[:, 1] - t_begin) / 1e3
print(f'  last launch: first CTA start -> last CTA end {span:.1f} us; {len(per_sm)} SMs used; co-resident CTAs per SM '
      f'max {max(peak)} / min {min(peak)}; CTA lifetime mean {life.mean():.1f} us max {life.max():.1f} us; '
      f'CTA start times: median {np.median(starts):.1f} us, 90 % {np.percentile(starts, 90):.1f} us, last {starts[-1]:.1f} us')
