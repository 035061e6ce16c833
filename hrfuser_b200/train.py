"""One CUDA graph for a whole training step of the backbone (configs[3]: HRFuser-B, SyncBN).

The reference trains through mmcv's `EpochBasedRunner` + `MMDistributedDataParallel`
(tools/train.py, mmdet/apis/train.py:135-160): ~11 000 kernel launches per step, issued one
by one from Python, the GPU idle between most of them (the step is host-bound on a B200:
profiles/r02_train_*.json).  The SyncBN statistics exchange adds one blocking collective per
norm layer and pass, the gradient exchange one bucketed all-reduce per ~25 MB.

`GraphedTrainStep` captures forward + backward + gradient exchange + optimizer step into ONE
CUDA graph over static input buffers:
  * the hrf_bn_* kernels (csrc/bn_train.cuh) and their fp64 statistics all-reduces
    (bn_train.py) are captured in place -- NCCL collectives are graph nodes, no host
    round trip per norm layer;
  * the gradients of all parameters that receive one are exchanged with ONE all-reduce over a
    flat bucket (sized for launch latency, not link count: NVSwitch gives every GPU full
    bandwidth to every peer), averaged over the ranks as DDP does; packed and unpacked with
    multi-tensor kernels (cat / _foreach_copy_), not one launch per parameter;
  * parameters that receive no gradient (unused transitions of a config) are found during the
    eager warm-up steps, not per step.
One process per GPU; `world == 1` captures the same graph without collectives.
"""
import torch
import torch.distributed as dist

from .bn_train import sync_group


def exchange_gradients(params, group, world):
    """Average the gradients of `params` over the ranks with ONE all-reduce of one flat bucket
    (parameters without a gradient are skipped; the set must be the same on every rank)."""
    grads = [p.grad for p in params if p.grad is not None]
    if world == 1 or not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world)
    # one multi-tensor copy back (a handful of launches), not one elementwise kernel per parameter:
    # HRFuser-B has ~1 500 parameter tensors, the step is bound by its kernel count
    torch._foreach_copy_(grads, [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in grads]), grads)])
    return flat.numel()


class GraphedTrainStep:
    """step = zero_grad -> loss_fn(net(x, mods)) -> backward -> all-reduce(grads) / world -> opt.step()

    x / mods are the static input buffers (copy new data into `self.x` / `self.mods`, or pass
    tensors to `__call__`).  `loss` holds the loss of the last replay."""

    def __init__(self, net, opt, x, mods, loss_fn, warmup=3, group=None):
        self.net, self.opt, self.loss_fn = net, opt, loss_fn
        self.x, self.mods = x, list(mods)
        self.group = sync_group(group)
        self.world = dist.get_world_size(self.group) if self.group is not None else 1
        self.params = [p for p in net.parameters() if p.requires_grad]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):          # allocator warm-up, momentum buffers, cuDNN plans
                self._step_eager()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = self._step_eager(zero=False)
        torch.cuda.synchronize()

    def _step_eager(self, zero=True):
        if zero:
            self.opt.zero_grad(set_to_none=True)
        loss = self.loss_fn(self.net(self.x, self.mods))
        loss.backward()
        exchange_gradients(self.params, self.group, self.world)
        self.opt.step()
        return loss.detach()

    def __call__(self, x=None, mods=None):
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if mods is not None:
            for d, m in zip(self.mods, mods):
                d.copy_(m, non_blocking=True)
        self.graph.replay()
        return self.loss
