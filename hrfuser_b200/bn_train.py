"""Train-mode BatchNorm2d / SyncBatchNorm on the hrf_bn_* kernels.

The reference builds its norms with mmcv `build_norm_layer` (`nn.BatchNorm2d` for
`type='BN'`, `nn.SyncBatchNorm` for `type='SyncBN'`; call sites hrnet.py:338-358,
hrformer.py:267-282, resnet.py:161-206, hrfuser_hrformer_based.py:380-397).  In training
each of them is (i) a per-channel reduction over (B, H, W), (ii) for SyncBN an exchange of
the statistics between the ranks, (iii) a per-channel affine pass; the backward is the
same three steps on (x, dy).  Here (i) and (iii) are the bandwidth-bound kernels of
`csrc/bn_train.cuh` and (ii) is ONE all-reduce of `2C + 1` fp64 values (sum, sum of
squares, count) over NCCL — the only data-path collective of the training configs
(torch's SyncBatchNorm all-gathers mean / invstd / count instead).

`HrfBatchNorm2d` / `HrfSyncBatchNorm` subclass the torch modules: same parameters,
buffers, `state_dict` keys and eval behaviour (eval BN is folded into the packed weights
by the engine and never runs here).  CUDA inputs in training mode take the kernel path —
there is no fallback for them: a missing library raises.  CPU tensors (the reference's own
CPU execution model, used by the CPU tests) go through torch's implementation.
"""
import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops


def _stats_from_sums(sums, count, eps):
    """fp64 [2C] (sum | sum of squares) and the element count -> fp64 mean, biased var, invstd."""
    C = sums.numel() // 2
    mean = sums[:C] / count
    var = (sums[C:] / count - mean * mean).clamp_min_(0.0)
    return mean, var, torch.rsqrt(var + eps)


def _all_reduce_sum(t, group):
    if group is not None:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class _BatchNormTrainFn(torch.autograd.Function):
    """y = BN_train(x); returns (y, mean, biased var, total count) — the last three are
    non-differentiable outputs used for the running statistics."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, group):
        x = x.contiguous()
        C = x.shape[1]
        local_n = x.numel() // C
        # [sum x | sum x^2 | count]: one fp64 message per BN for SyncBN
        msg = torch.empty(2 * C + 1, dtype=torch.float64, device=x.device)
        msg[:2 * C] = ops.bn_stats(x)
        msg[2 * C] = local_n
        _all_reduce_sum(msg, group)
        count = msg[2 * C]
        mean, var, invstd = _stats_from_sums(msg[:2 * C], count, eps)
        w = weight.double() if weight is not None else torch.ones_like(mean)
        b = bias.double() if bias is not None else torch.zeros_like(mean)
        a = w * invstd
        y = ops.bn_affine(x, a.float(), (b - mean * a).float())
        mean32, var32, invstd32 = mean.float(), var.float(), invstd.float()
        ctx.save_for_backward(x, weight, mean32, invstd32, count)
        ctx.group = group
        ctx.mark_non_differentiable(mean32, var32, count)
        return y, mean32, var32, count

    @staticmethod
    def backward(ctx, dy, _dmean, _dvar, _dcount):
        x, weight, mean, invstd, count = ctx.saved_tensors
        dy = dy.contiguous()
        C = x.shape[1]
        sums = ops.bn_bwd_stats(x, dy, mean, invstd)       # local sum dy | sum dy * xhat
        # parameter gradients are rank-local (DDP averages them), as in torch's SyncBatchNorm
        dweight = sums[C:].to(weight.dtype) if weight is not None and ctx.needs_input_grad[1] else None
        dbias = sums[:C].to(dy.dtype) if ctx.needs_input_grad[2] else None
        dx = None
        if ctx.needs_input_grad[0]:
            tot = _all_reduce_sum(sums.clone(), ctx.group) if ctx.group is not None else sums
            mdy, mdyx = tot[:C] / count, tot[C:] / count
            g = (weight.double() if weight is not None else torch.ones_like(mdy)) * invstd.double()
            # dx = g * (dy - mean(dy) - xhat * mean(dy * xhat)),  xhat = (x - mean) * invstd
            kb = -g * invstd.double() * mdyx
            kc = -g * mdy - kb * mean.double()
            dx = ops.bn_affine(x, g.float(), kc.float(), dy=dy, b=kb.float())
        return dx, dweight, dbias, None, None


def _train_forward(mod, x, group):
    if mod.momentum is None:
        raise NotImplementedError('cumulative moving average (momentum=None) is not supported')
    y, mean, var, count = _BatchNormTrainFn.apply(x, mod.weight, mod.bias, mod.eps, group)
    if mod.track_running_stats and mod.running_mean is not None:
        with torch.no_grad():
            m = mod.momentum
            unbiased = var * (count / (count - 1).clamp_min(1.0)).float()
            mod.running_mean.mul_(1 - m).add_(mean.to(mod.running_mean.dtype), alpha=m)
            mod.running_var.mul_(1 - m).add_(unbiased.to(mod.running_var.dtype), alpha=m)
            mod.num_batches_tracked += 1
    return y


class HrfBatchNorm2d(nn.BatchNorm2d):
    """`nn.BatchNorm2d` whose CUDA training forward / backward run on hrf_bn_* kernels."""

    def forward(self, x):
        if x.is_cuda and (self.training or not self.track_running_stats):
            self._check_input_dim(x)
            return _train_forward(self, x, None)
        return super().forward(x)


class HrfSyncBatchNorm(nn.SyncBatchNorm):
    """`nn.SyncBatchNorm` with one fp64 all-reduce of (sum, sum of squares, count) per forward
    and one of (sum dy, sum dy * xhat) per backward, around the hrf_bn_* kernels."""

    def forward(self, x):
        if x.is_cuda and (self.training or not self.track_running_stats):
            self._check_input_dim(x)
            return _train_forward(self, x, sync_group(self.process_group))
        return super().forward(x)


def sync_group(process_group=None):
    """The group to all-reduce over, or None when there is nothing to synchronise with."""
    if not (dist.is_available() and dist.is_initialized()):
        return None
    group = process_group if process_group is not None else dist.group.WORLD
    return group if dist.get_world_size(group) > 1 else None
