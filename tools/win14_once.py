"""One LSA call per HRFuser-T width at window 14 (bf16), for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/win14_once.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import make_block  # noqa: E402
from hrfuser_b200 import ops  # noqa: E402
from hrfuser_b200.engine import BackboneEngine  # noqa: E402

win = int(os.environ.get('WIN', '14'))
for (H, W), (C, heads) in zip([(96, 160), (48, 80), (24, 40), (12, 20)], [(18, 1), (36, 2), (72, 4), (144, 8)]):
    e = BackboneEngine.__new__(BackboneEngine)
    e._host_blobs, e._blob_slots = [], []
    e.device, e.dtype = torch.device('cuda'), torch.float32
    blk, _ = make_block('lsa', C, heads, win=win)
    pk = e._hrformer_block(blk)
    e._upload()
    x = torch.randn(8, H, W, C, device='cuda').bfloat16()
    for _ in range(3):
        ops.window_attention(x, None, [s.t for s in pk['attn']], heads, win=win)
    torch.cuda.synchronize()
