"""SyncBN over NCCL: HrfSyncBatchNorm (hrf_bn_* kernels + one fp64 all-reduce per pass) against
torch.nn.SyncBatchNorm on the same shards, then a timed HRFuser training step (BASELINE.json
configs[3]: HRFuser-B nuScenes, SyncBN, 2 frames per GPU; loss = sum of squared outputs) with
either BN implementation.  Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/syncbn_nccl_check.py [--variant b] [--steps 5]
"""
import argparse
import copy
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hrfuser_b200 import HRFuserHRFormerBased, backbone_cfg, bn_train, ops  # noqa: E402
from hrfuser_b200 import dist as hdist  # noqa: E402
from hrfuser_b200.utils import randomize_parameters, synthetic_inputs  # noqa: E402


def module_parity(rank, world):
    worst = {}
    for shape in [(2, 64, 192, 320), (2, 78, 96, 160), (3, 144, 12, 20)]:
        g = torch.Generator().manual_seed(100 + rank)
        B = shape[0] + rank                                   # ragged shards
        x = (torch.randn(B, *shape[1:], generator=g) * 2 + 1).cuda()
        dy = torch.randn(B, *shape[1:], generator=g).cuda()
        mods = [bn_train.HrfSyncBatchNorm(shape[1]).cuda().train(),
                nn.SyncBatchNorm(shape[1]).cuda().train()]
        res = []
        for m in mods:
            with torch.no_grad():
                m.weight.copy_(torch.linspace(0.5, 1.5, shape[1]))
                m.bias.copy_(torch.linspace(-1, 1, shape[1]))
            xr = x.clone().requires_grad_(True)
            y = m(xr)
            y.backward(dy)
            res.append(dict(y=y.detach(), dx=xr.grad, dw=m.weight.grad, db=m.bias.grad,
                            rm=m.running_mean, rv=m.running_var))
        for k in res[0]:
            e = float((res[0][k] - res[1][k]).norm() / (res[1][k].norm() + 1e-12))
            worst[k] = max(worst.get(k, 0.0), e)
    t = torch.tensor([worst[k] for k in sorted(worst)], device='cuda', dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return dict(zip(sorted(worst), t.tolist()))


def train_step_ms(net, x, mods, steps, warmup):
    def step():
        net.zero_grad(set_to_none=True)
        out = net(x, mods)
        sum((o * o).mean() for o in out).backward()
    for _ in range(warmup):
        step()
    hdist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return hdist.max_over_ranks(e0.elapsed_time(e1) / steps, torch.device('cuda'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--variant', default='b')
    ap.add_argument('--frames', type=int, default=2)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=2)
    a = ap.parse_args()
    rank, world, local = hdist.init_from_env('nccl')
    torch.cuda.set_device(local)
    parity = module_parity(rank, world)

    c = backbone_cfg(a.variant, 'nus')
    c.pop('type')
    c['norm_cfg'] = dict(type='SyncBN', requires_grad=True)
    net = HRFuserHRFormerBased(**c)
    randomize_parameters(net, 1)
    net = net.cuda().train()
    ref = copy.deepcopy(net)
    for m in ref.modules():
        if isinstance(m, bn_train.HrfSyncBatchNorm):
            m.__class__ = nn.SyncBatchNorm
    x, mods = synthetic_inputs(a.frames, 384, 640, (3, 3), seed=10 + rank)
    x, mods = x.cuda(), [m.cuda() for m in mods]
    lib = ops._lib.load()
    n0 = lib.hrf_launch_count()
    ms_k = train_step_ms(net, x, mods, a.steps, a.warmup)
    launches = (lib.hrf_launch_count() - n0) // (a.steps + a.warmup)
    ms_t = train_step_ms(ref, x, mods, a.steps, a.warmup)
    if rank == 0:
        print(json.dumps(dict(
            what='SyncBN training step, forward + backward, torch autograd around the BN kernels',
            model=f'hrfuser_{a.variant}_nus_r640', n_gpus=world, frames_per_gpu=a.frames,
            ms_per_step_hrf_syncbn=round(ms_k, 2), ms_per_step_torch_syncbn=round(ms_t, 2),
            frames_per_s_hrf_syncbn=round(world * a.frames / ms_k * 1e3, 1),
            frames_per_s_torch_syncbn=round(world * a.frames / ms_t * 1e3, 1),
            hrf_bn_launches_per_step=int(launches),
            module_parity_vs_torch_syncbn_rel_l2=parity)))
    hdist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
