"""Isolated kernel micro-benchmark (BASELINE.json configs[4]): window attention (LSA, MWCA
with 1-3 modalities) and MixFFN over every HRNet resolution of the nuScenes / STF grids.

    python tools/microbench.py [--batch 8] [--precision bf16] [--iters 30] [--out file.jsonl]

Each shape: the op is launched `iters` times back to back between two CUDA events (the GPU
is never starved by the host), inputs rotate over enough buffers to exceed the 126 MB L2.
Prints one JSON line per shape with the measured time and the roofline fractions against
MEASURED_PEAKS.json (algorithmic bytes / FLOPs of SURVEY.md section 8d).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from bench import peaks  # noqa: E402
from helpers import make_block  # noqa: E402
from hrfuser_b200 import ops  # noqa: E402
from hrfuser_b200.engine import BackboneEngine  # noqa: E402

GRIDS = {'nus': [(96, 160), (48, 80), (24, 40), (12, 20)],
         'stf': [(96, 312), (48, 156), (24, 78), (12, 39)]}
WIDTHS = [(18, 1), (36, 2), (72, 4), (144, 8)]                  # HRFuser-T: head_dim 18
WIDTHS_B = [(78, 2), (156, 4), (312, 8), (624, 16)]             # HRFuser-B: head_dim 39


def stub():
    e = BackboneEngine.__new__(BackboneEngine)
    e._host_blobs, e._blob_slots = [], []
    e.device, e.dtype = torch.device('cuda'), torch.float32
    return e


def timed(fn, n_sets, iters):
    """Average time of one call: `iters` calls captured into ONE CUDA graph (the host launch path,
    15-25 us per call through ctypes, is then out of the measurement), replayed three times."""
    for i in range(3):
        fn(i % n_sets)
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        fn(0)
        st.synchronize()
        with torch.cuda.graph(g, stream=st):
            for i in range(iters):
                fn(i % n_sets)
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / iters)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--precision', default='bf16')
    ap.add_argument('--iters', type=int, default=30)
    ap.add_argument('--grids', default='nus,stf')
    ap.add_argument('--wins', default='7', help='window sizes, e.g. 7,14')
    ap.add_argument('--out', default='')
    ap.add_argument('--kinds', default='lsa,mwca,mixffn', help='subset of lsa,mwca,mixffn')
    ap.add_argument('--widths', default='', help='channel widths to run (default: all of the variant)')
    ap.add_argument('--variant', default='t', choices=['t', 'b'], help='HRFuser-T (head_dim 18) or -B (39)')
    ap.add_argument('--mods', default='1,2,3', help='MWCA modality counts')
    a = ap.parse_args()
    dt = torch.bfloat16 if a.precision == 'bf16' else torch.float32
    pk = peaks()
    ridge = pk['bf16_tflops'] * 1e12 / (pk['hbm_gbs'] * 1e9)
    out = open(a.out, 'w') if a.out else None
    B = a.batch
    for gname, win in [(g, int(w)) for w in a.wins.split(',') for g in a.grids.split(',')]:
        S4 = 4 * win * win                                   # QK^T + PV flops per token and channel
        for (H, W), (C, heads) in zip(GRIDS[gname], WIDTHS if a.variant == 't' else WIDTHS_B):
            if a.widths and str(C) not in a.widths.split(','):
                continue
            n_tok = B * H * W
            nbytes = n_tok * C * dt.itemsize
            n_sets = max(2, int(2 * 126e6 // max(nbytes, 1)) + 1)
            n_sets = min(n_sets, 64)
            xs = [torch.randn(B, H, W, C, device='cuda').to(dt) for _ in range(n_sets)]
            cases = [('lsa', 0), ('mwca', 1), ('mwca', 2), ('mwca', 3)] + ([('mixffn', 0)] if win == 7 else [])
            cases = [c for c in cases if c[0] in a.kinds.split(',') and (c[0] != 'mwca' or str(c[1]) in a.mods.split(','))]
            for kind, M in cases:
                e = stub()
                if kind == 'mixffn':
                    blk, _ = make_block('lsa', C, heads)
                    f = e._ffn(blk.norm2, blk.ffn)
                    e._upload()
                    fn = lambda i: ops.mixffn(xs[i], f['blob'].t, f['hidden'], f['eps'])
                    flops, byts = n_tok * (16 * C * C + 72 * C), 2 * nbytes
                elif kind == 'lsa':
                    blk, _ = make_block('lsa', C, heads, win=win)
                    pk_ = e._hrformer_block(blk)
                    e._upload()
                    blobs = [s.t for s in pk_['attn']]
                    fn = lambda i: ops.window_attention(xs[i], None, blobs, heads, win=win)
                    flops, byts = n_tok * (8 * C * C + S4 * C), 2 * nbytes
                else:
                    blk, _ = make_block('mwca', C, heads, M=M, win=win)
                    pk_ = e._fusion_block(blk)
                    e._upload()
                    blobs = [s.t for s in pk_['attn']]
                    zs = [xs[(k + 1) % n_sets] for k in range(M)]
                    fn = lambda i: ops.window_attention(xs[i], zs, blobs, heads, win=win)
                    flops, byts = M * n_tok * (8 * C * C + S4 * C), (2 + M) * nbytes
                ms = timed(fn, n_sets, a.iters)
                gbs, tfs = byts / ms / 1e6, flops / ms / 1e9
                bound = 'hbm' if flops / byts < ridge else 'tensor'
                rec = dict(kind=kind, modalities=M, grid=gname, B=B, H=H, W=W, C=C, heads=heads,
                           win=win, precision=a.precision, ms=round(ms, 5), GBps=round(gbs, 1),
                           TFLOPs=round(tfs, 2), bound=bound,
                           frac_hbm=round(gbs / pk['hbm_gbs'], 4),
                           frac_tensor=round(tfs / pk['bf16_tflops'], 5))
                line = json.dumps(rec)
                print(line, flush=True)
                if out:
                    out.write(line + '\n')
    if out:
        out.close()


if __name__ == '__main__':
    main()
