"""One un-graphed backbone step between cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none \
      --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on \
      -k regex:mixffn -c 3 -o gpurun_out/prof_ffn python tools/profile_step.py
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_net  # noqa: E402
from hrfuser_b200.utils import synthetic_inputs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='hrfuser_t_nus_r640')
ap.add_argument('--batch', type=int, default=8)
ap.add_argument('--precision', default='bf16')
a = ap.parse_args()
dev = torch.device('cuda', 0)
cfg, net, (H, W), mod_ch = build_net(a.workload, a.precision, dev)
x, mods = synthetic_inputs(a.batch, H, W, mod_ch, seed=0, device=dev)
with torch.no_grad():
    for _ in range(2):
        net(x, mods)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    net(x, mods)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
