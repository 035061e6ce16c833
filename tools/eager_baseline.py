"""PyTorch-eager numbers on the GPU box, for context.

The unmodified reference cannot travel to the GPU box (Python + mmcv, sources may not be
copied), but `hrfuser_b200/modules.py` is the same sequence of torch ops, verified against the
reference to 2e-7 (tests/test_engine_cpu.py::test_training_path_matches_reference_golden).
This script times that eager path -- what the reference's own forward costs on a B200 --
next to the engine, and one training step (forward + backward + AdamW) for config 4.

    python tools/eager_baseline.py [--batch 8] [--workload hrfuser_t_nus_r640]
    torchrun --nproc-per-node 2 tools/eager_baseline.py --train --workload hrfuser_b_nus_r640 --syncbn
"""
import argparse
import copy
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hrfuser_b200 import WORKLOADS, HRFuserHRFormerBased, backbone_cfg  # noqa: E402
from hrfuser_b200 import dist as hdist  # noqa: E402
from hrfuser_b200.utils import randomize_parameters, synthetic_inputs  # noqa: E402


def timeit(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='hrfuser_t_nus_r640')
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--train', action='store_true')
    ap.add_argument('--syncbn', action='store_true')
    a = ap.parse_args()
    rank, world, local = hdist.init_from_env()
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    variant, dataset, H, W = WORKLOADS[a.workload]
    cfg = backbone_cfg(variant, dataset, norm='SyncBN' if a.syncbn else 'BN')
    c = copy.deepcopy(cfg)
    c.pop('type')
    mod_ch = tuple(c.get('mod_in_channels', [3, 3]))
    net = HRFuserHRFormerBased(**c, precision='bf16')
    randomize_parameters(net, 0)
    net.to(dev)
    x, mods = synthetic_inputs(a.batch, H, W, mod_ch, seed=rank, device=dev)
    out = dict(workload=a.workload, batch_per_gpu=a.batch, n_gpus=world)
    if a.train:
        # config 4: training step with batch statistics (SyncBN over NCCL when --syncbn),
        # DropPath / Dropout active, loss = sum of the outputs, AdamW
        net.train()
        model = net
        if world > 1:
            model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local],
                                                              find_unused_parameters=True)
        opt = torch.optim.AdamW(net.parameters(), lr=3e-4, weight_decay=0.01)

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast('cuda', dtype=torch.bfloat16):
                ys = model(x, [m.clone() for m in mods])
            sum(y.float().mean() for y in ys).backward()
            opt.step()
        ms = hdist.max_over_ranks(timeit(step, a.iters), dev)
        out.update(mode='train step (torch autograd path, bf16 autocast)', ms_per_step=ms,
                   frames_per_s=a.batch * world / ms * 1e3, norm='SyncBN' if a.syncbn else 'BN')
    else:
        net.eval()
        with torch.no_grad():
            eager32 = timeit(lambda: net._forward_autograd(x, list(mods)), a.iters)
            with torch.autocast('cuda', dtype=torch.bfloat16):
                eager16 = timeit(lambda: net._forward_autograd(x, list(mods)), a.iters)
            eng = timeit(lambda: net(x, mods), a.iters)
        out.update(mode='eval forward', eager_fp32_ms=eager32, eager_bf16_autocast_ms=eager16,
                   engine_bf16_eager_launch_ms=eng,
                   eager_fp32_fps=a.batch / eager32 * 1e3, eager_bf16_fps=a.batch / eager16 * 1e3,
                   engine_fps=a.batch / eng * 1e3)
    if rank == 0:
        print(json.dumps(out))
    hdist.barrier()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
