"""Per-phase cycle breakdown of the tcgen05 MixFFN kernel (instrumented build).

    nvcc ... -DHRF_FFN_PROFILE -o gpurun_out/libhrf_prof.so hrfuser_b200/csrc/abi.cu   (tools/gpu_ffn_phases.sh)
    HRF_LIB=gpurun_out/libhrf_prof.so python tools/ffn_phases.py

Thread 0 of every CTA accumulates clock64 deltas per phase; this prints the mean cycles per
tile of each phase over all CTAs (and the per-CTA setup time).
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import make_block  # noqa: E402
from hrfuser_b200 import _lib, ops  # noqa: E402
from microbench import stub  # noqa: E402

NAMES = ['LN prologue', 'sync', 'fc1 issue', 'fc1 wait', 'epilogue 1', 'sync', 'conv issue',
         'conv wait', 'epilogue dw', 'sync', 'fc2 issue', 'fc2 wait', 'epilogue 2', '-', 'setup', 'tiles']
lib = _lib.load()
lib.hrf_debug_ffn_prof.argtypes = [C.c_void_p, C.c_int, C.c_int]
Cc, heads, B, H, W = 18, 1, 8, 96, 160
blk, _ = make_block('lsa', Cc, heads)
eng = stub()
f = eng._ffn(blk.norm2, blk.ffn)
eng._upload()
blob = f['blob'].t
x = torch.randn(B, H, W, Cc, device='cuda').to(torch.bfloat16)
for it in range(3):
    ops.mixffn(x, blob, 4 * Cc)
torch.cuda.synchronize()
lib.hrf_debug_ffn_prof(None, 0, 1)
n_it = 10
for it in range(n_it):
    ops.mixffn(x, blob, 4 * Cc)
torch.cuda.synchronize()
buf = np.zeros(1024 * 16, dtype=np.uint64)
lib.hrf_debug_ffn_prof(buf.ctypes.data, buf.size, 0)
a = buf.reshape(1024, 16).astype(np.float64)
a = a[a[:, 15] > 0]
tiles = a[:, 15].sum()
print(f'{len(a)} CTAs, {tiles / n_it:.0f} tiles per launch, {a[:, 15].mean() / n_it:.2f} tiles per CTA')
tot = 0.0
for k in range(13):
    v = a[:, k].sum() / tiles
    tot += v
    print(f'  {NAMES[k]:12s} {v:9.0f} cycles / tile')
print(f'  {"total":12s} {tot:9.0f} cycles / tile;  setup {a[:, 14].mean() / n_it:.0f} cycles / CTA')
