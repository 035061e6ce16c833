"""Stem + layer1 + transition1 chain of the camera stream: whole batch vs batch chunks (so the
256-channel intermediates stay L2-resident between producer and consumer).  Graph-timed."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_net  # noqa: E402
from hrfuser_b200.utils import synthetic_inputs  # noqa: E402

dev = torch.device('cuda', 0)
cfg, net, (H, W), mod_ch = build_net('hrfuser_t_nus_r640', 'bf16', dev)
eng = net.engine()
x, mods = synthetic_inputs(8, H, W, mod_ch, seed=0, device=dev)


def chain(xb):
    y = eng._apply_chain(eng.stem, xb)
    return [t(y) for t in eng.trans1]


def graph_time(fn, n=5):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s), torch.no_grad():
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5 / n * 1e3


for nchunk in (1, 2, 4, 8):
    def run():
        outs = [chain(xc) for xc in x.chunk(nchunk)]
        return [torch.cat([o[i] for o in outs]) for i in range(2)] if nchunk > 1 else outs[0]
    print(f'camera stem chain, {nchunk} batch chunk(s) of {8 // nchunk} frames: {graph_time(run):8.1f} us', flush=True)
