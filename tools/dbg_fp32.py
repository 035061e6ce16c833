import os, sys, torch
sys.path.insert(0, '/root/repo')
from bench import build_net
from hrfuser_b200.utils import synthetic_inputs
from hrfuser_b200 import ops
dev = torch.device('cuda', 0)
cfg, net, (H, W), mod_ch = build_net('hrfuser_t_nus_r640', 'fp32', dev)
for B in (2, 8):
    x, mods = synthetic_inputs(B, H, W, mod_ch, seed=0, device=dev)
    eng = net.engine()
    for conc in (False, True):
        eng.concurrent = conc
        with torch.no_grad():
            out = net(x, mods)
        torch.cuda.synchronize()
        print('B', B, 'concurrent', conc, 'ok', [float(o.abs().mean()) for o in out], flush=True)
