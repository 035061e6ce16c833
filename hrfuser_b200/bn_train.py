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
import torch.nn.functional as F

from . import ops


def _all_reduce_sum(t, group):
    if group is not None:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class _BatchNormTrainFn(torch.autograd.Function):
    """y = BN_train(x).  Forward: hrf_bn_stats -> [all-reduce] -> hrf_bn_normalize (which also
    saves mean / invstd and updates the running statistics); backward: hrf_bn_bwd_stats ->
    [all-reduce] -> hrf_bn_bwd_dx.  No per-channel math on the host side.
    `act` fuses the ReLU / GELU module that follows the norm layer: y = act(BN_train(x)); the
    backward recomputes the pre-activation from x, so only x is saved."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, group, act=0):
        x = x.contiguous()
        stats = _all_reduce_sum(ops.bn_stats(x), group)      # [sum x | sum x^2 | count], fp64
        y, mean, invstd = ops.bn_normalize(x, stats, weight, bias, eps, momentum, running_mean,
                                           running_var, act=act)
        ctx.save_for_backward(x, weight, bias, mean, invstd, stats)
        ctx.group, ctx.act = group, act
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias, mean, invstd, stats = ctx.saved_tensors
        dy = dy.contiguous()
        C = x.shape[1]
        # parameter gradients are rank-local (DDP averages them), as in torch's SyncBatchNorm
        sums, dweight, dbias = ops.bn_bwd_stats(x, dy, mean, invstd, want_param_grads=True,
                                                weight=weight, bias=bias, act=ctx.act)
        dx = None
        if ctx.needs_input_grad[0]:
            _all_reduce_sum(sums, ctx.group)
            dx = ops.bn_bwd_dx(x, dy, sums, stats[2 * C:], weight, mean, invstd, bias=bias,
                               act=ctx.act)
        return (dx, dweight if weight is not None and ctx.needs_input_grad[1] else None,
                dbias if bias is not None and ctx.needs_input_grad[2] else None,
                None, None, None, None, None, None)


def _train_forward(mod, x, group, act=0):
    if mod.momentum is None:
        raise NotImplementedError('cumulative moving average (momentum=None) is not supported')
    track = mod.track_running_stats and mod.running_mean is not None
    y = _BatchNormTrainFn.apply(x, mod.weight, mod.bias, mod.running_mean if track else None,
                                mod.running_var if track else None, mod.eps, mod.momentum, group, act)
    if track:
        mod.num_batches_tracked.add_(1)
    return y


class HrfBatchNorm2d(nn.BatchNorm2d):
    """`nn.BatchNorm2d` whose CUDA training forward / backward run on hrf_bn_* kernels."""

    def forward(self, x, act=0):
        if x.is_cuda and (self.training or not self.track_running_stats):
            self._check_input_dim(x)
            return _train_forward(self, x, None, act)
        return _apply_act(super().forward(x), act)


class HrfSyncBatchNorm(nn.SyncBatchNorm):
    """`nn.SyncBatchNorm` with one fp64 all-reduce of (sum, sum of squares, count) per forward
    and one of (sum dy, sum dy * xhat) per backward, around the hrf_bn_* kernels."""

    def forward(self, x, act=0):
        if x.is_cuda and (self.training or not self.track_running_stats):
            self._check_input_dim(x)
            return _train_forward(self, x, sync_group(self.process_group), act)
        return _apply_act(super().forward(x), act)


def _apply_act(y, act):
    return F.relu(y) if act == ops.BN_ACT_RELU else F.gelu(y) if act == ops.BN_ACT_GELU else y


def act_code(m):
    """BN_ACT_* of an activation module the BN kernels can fuse, else None"""
    if isinstance(m, nn.ReLU):
        return ops.BN_ACT_RELU
    if isinstance(m, nn.GELU) and getattr(m, 'approximate', 'none') == 'none':
        return ops.BN_ACT_GELU
    return None


def norm_act(norm, act, x):
    """act(norm(x)), in the norm layer's own kernels when it is an Hrf(Sync)BatchNorm2d on CUDA"""
    code = act_code(act) if isinstance(norm, (HrfBatchNorm2d, HrfSyncBatchNorm)) and x.is_cuda else None
    if code is not None and NormActSequential.fuse_act:
        return norm(x, act=code)
    return act(norm(x))


class NormActSequential(nn.Sequential):
    """nn.Sequential (same child names, same state_dict) that hands an activation module to the
    Hrf(Sync)BatchNorm2d right before it, whose kernels apply it and its derivative in their
    own passes (`fuse_act = False` restores the child-by-child walk)."""
    fuse_act = True

    def forward(self, x):
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            code = act_code(nxt) if self.fuse_act and isinstance(m, (HrfBatchNorm2d, HrfSyncBatchNorm)) else None
            if code is not None and x.is_cuda:
                x = m(x, act=code)
                i += 2
            else:
                x = m(x)
                i += 1
        return x


def sync_group(process_group=None):
    """The group to all-reduce over, or None when there is nothing to synchronise with."""
    if not (dist.is_available() and dist.is_initialized()):
        return None
    group = process_group if process_group is not None else dist.group.WORLD
    return group if dist.get_world_size(group) > 1 else None


class _LayerNormTrainFn(torch.autograd.Function):
    """y = LayerNorm(x) over the last axis on hrf_ln_fwd / hrf_ln_bwd (csrc/ln_train.cuh)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        x = x.contiguous()
        y, mean, rstd = ops.ln_fwd(x, weight, bias, eps)
        ctx.save_for_backward(x, weight, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, mean, rstd = ctx.saved_tensors
        dx, dw, db = ops.ln_bwd(x, dy.contiguous(), mean, rstd, weight, want_dx=ctx.needs_input_grad[0])
        return dx, dw if ctx.needs_input_grad[1] else None, db if ctx.needs_input_grad[2] else None, None


class HrfLayerNorm(nn.LayerNorm):
    """`nn.LayerNorm` over the channel axis whose CUDA fp32 forward / backward under autograd run
    on the hrf_ln_* kernels (the training path; the inference engine folds LayerNorm into the
    attention / MixFFN kernels and never calls this).  Same parameters and state_dict keys."""

    def forward(self, x):
        if (x.is_cuda and x.dtype == torch.float32 and len(self.normalized_shape) == 1 and
                self.elementwise_affine and self.bias is not None and self.normalized_shape[0] <= 1024 and
                torch.is_grad_enabled()):
            return _LayerNormTrainFn.apply(x, self.weight, self.bias, self.eps)
        return super().forward(x)


class _DepthwiseConvTrainFn(torch.autograd.Function):
    """depthwise 3x3 conv (pad 1, stride 1 | 2) on hrf_dwconv_train_* (csrc/dwconv_train.cuh)"""

    @staticmethod
    def forward(ctx, x, weight, bias, stride):
        x = x.contiguous()
        ctx.save_for_backward(x, weight)
        ctx.stride, ctx.has_bias = stride, bias is not None
        return ops.dwconv_train_fwd(x, weight, bias, stride)

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        dx, dw, db = ops.dwconv_train_bwd(x, g.contiguous(), weight, ctx.stride, want_dx=ctx.needs_input_grad[0],
                                          want_dw=ctx.needs_input_grad[1],
                                          want_dbias=ctx.has_bias and ctx.needs_input_grad[2])
        return dx, dw if ctx.needs_input_grad[1] else None, db, None


class HrfDepthwiseConv2d(nn.Conv2d):
    """`nn.Conv2d(C, C, 3, stride, 1, groups=C)` whose CUDA fp32 forward / backward under autograd
    run on the hrf_dwconv_train_* kernels; anything else (CPU, other dtypes, no-grad) is torch's."""

    use_kernels = True

    def forward(self, x):
        if (self.use_kernels and x.is_cuda and x.dtype == torch.float32 and torch.is_grad_enabled() and
                x.dim() == 4 and self.weight.is_contiguous()):
            return _DepthwiseConvTrainFn.apply(x, self.weight, self.bias, self.stride[0])
        return super().forward(x)


def make_conv(cin, cout, k, stride=1, padding=0, groups=1, bias=True):
    """nn.Conv2d, or its depthwise-3x3 subclass on the library's training kernels"""
    if groups == cin == cout and k == 3 and padding == 1 and stride in (1, 2):
        return HrfDepthwiseConv2d(cin, cout, 3, stride, 1, groups=groups, bias=bias)
    return nn.Conv2d(cin, cout, k, stride, padding, groups=groups, bias=bias)
