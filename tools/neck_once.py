"""one HRFPN.reduce call per precision mode at the HRFuser-T nuScenes size (for ncu launch lists)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hrfuser_b200.neck import HRFPN  # noqa: E402

chans = [18, 36, 72, 144]
xs = [torch.randn(8, c, 96 >> i, 160 >> i, device='cuda') for i, c in enumerate(chans)]
for mode in ('bf16', 'fp32'):
    net = HRFPN(in_channels=chans, out_channels=256, precision=mode).eval().cuda()
    with torch.no_grad():
        net.reduce(xs)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        net.reduce(xs)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
