"""Parameter containers for the HRFuser fusion backbone.

The classes below own the parameters/buffers under exactly the names the
reference module tree produces (reference
mmdet/models/backbones/hrfuser_hrformer_based.py:337-468, hrformer.py:38-94,
256-282,324-363,423-561, hrnet.py:288-510, resnet.py:97-206), so a reference
checkpoint loads key-for-key.  Each container also carries a small autograd-
capable `forward` written with torch ops: that is the *training-mode* path
(batch-statistics BN / SyncBN, DropPath, gradients).  Inference (`eval()`)
never runs these forwards -- `backbone.HRFuserHRFormerBased.forward` hands the
parameters to the sm_100a engine (`engine.py`) instead and raises if the CUDA
library is missing.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .bn_train import (HrfBatchNorm2d, HrfLayerNorm, HrfSyncBatchNorm, NormActSequential, make_conv,
                       norm_act)
from .window_maps import relative_position_index, window_geometry


# ----------------------------------------------------------------------------
# factories (the subset of mmcv.cnn builders the backbone relies on)
# ----------------------------------------------------------------------------
def make_norm(cfg, channels):
    cfg = dict(cfg)
    kind = cfg.pop('type')
    trainable = cfg.pop('requires_grad', True)
    cfg.setdefault('eps', 1e-5)
    if kind in ('BN', 'BN2d'):
        m = HrfBatchNorm2d(channels, **cfg)
    elif kind == 'SyncBN':
        m = HrfSyncBatchNorm(channels, **cfg)
    elif kind == 'LN':
        m = HrfLayerNorm(channels, **cfg)
    else:
        raise KeyError(f'unsupported norm type {kind!r}')
    for p in m.parameters():
        p.requires_grad = trainable
    return m


def conv_bn(cin, cout, k, stride, norm_cfg, relu, groups=1):
    layers = [make_conv(cin, cout, k, stride, k // 2, groups=groups, bias=False),
              make_norm(norm_cfg, cout)]
    if relu:
        layers.append(nn.ReLU(inplace=relu == 'inplace'))
    return NormActSequential(*layers)


class DropPath(nn.Module):
    """Per-sample stochastic depth (mmcv DropPath semantics)."""

    def __init__(self, p):
        super().__init__()
        self.drop_prob = p

    def forward(self, x):
        if not self.training or self.drop_prob == 0.:
            return x
        keep = 1. - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).uniform_().add_(keep).floor_()
        return x / keep * mask


def _tokens(x):
    """(B,C,H,W) -> (B,H*W,C)"""
    return x.flatten(2).transpose(1, 2)


def _image(x, H, W):
    """(B,H*W,C) -> (B,C,H,W)"""
    return x.transpose(1, 2).reshape(x.shape[0], x.shape[2], H, W)


# ----------------------------------------------------------------------------
# window attention
# ----------------------------------------------------------------------------
class _AttnCoreTrainFn(torch.autograd.Function):
    """softmax(scale q k^T + rel-pos bias) v per (window, head) on hrf_attn_core_train_fwd / _bwd
    (csrc/attn_train.cuh): q / k / v [nWin, N, C] with heads along C, as the Linear layers emit."""

    @staticmethod
    def forward(ctx, q, k, v, table, rpi, heads, scale):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        o, P = ops.attn_core_train_fwd(q, k, v, heads, scale, table, rpi)
        ctx.save_for_backward(q, k, v, P, rpi)
        ctx.heads, ctx.scale = heads, scale
        ctx.T = table.shape[0] if table is not None else 0
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, P, rpi = ctx.saved_tensors
        want_table = ctx.T > 0 and ctx.needs_input_grad[3]
        dq, dk, dv, dtable = ops.attn_core_train_bwd(q, k, v, P, do.contiguous(), ctx.heads, ctx.scale,
                                                     rpi if want_table else None, ctx.T)
        return dq, dk, dv, dtable, None, None, None


class _WindowAttnBase(nn.Module):
    """Shared pieces of WindowMSA / WindowMCA: rpb table + index, softmax core."""

    use_kernels = True      # False: the torch formulation everywhere (bench.py's torch arm, A/B runs)

    def _init_common(self, dim, heads, window, with_rpe):
        self.num_heads = heads
        self.Wh, self.Ww = window
        self.scale = (dim // heads) ** -0.5
        self.with_rpe = with_rpe
        if with_rpe:
            self.relative_position_bias_table = nn.Parameter(
                torch.zeros((2 * self.Wh - 1) * (2 * self.Ww - 1), heads))
            self.register_buffer('relative_position_index',
                                 torch.from_numpy(relative_position_index(self.Wh, self.Ww)))

    def _core(self, q, k, v, mask):
        nWin, N, C = q.shape
        h = self.num_heads
        if (self.use_kernels and q.is_cuda and q.dtype == torch.float32 and mask is None and torch.is_grad_enabled() and
                not (self.training and self.attn_drop.p > 0) and ops.attn_core_train_supported(N, C // h)):
            # training path on the GPU: the whole core (and its backward) is one kernel per pass
            if self.with_rpe and getattr(self, '_rpi32', None) is None or \
                    self.with_rpe and self._rpi32.device != q.device:
                self._rpi32 = self.relative_position_index.reshape(-1).to(device=q.device, dtype=torch.int32)
            o = _AttnCoreTrainFn.apply(q, k, v, self.relative_position_bias_table if self.with_rpe else None,
                                       self._rpi32 if self.with_rpe else None, h, self.scale)
            return self.proj_drop(self.out_proj(o))
        split = lambda t: t.view(nWin, N, h, C // h).transpose(1, 2)
        logits = (split(q) * self.scale) @ split(k).transpose(-1, -2)
        if self.with_rpe:
            bias = self.relative_position_bias_table[self.relative_position_index.reshape(-1)]
            logits = logits + bias.view(N, N, h).permute(2, 0, 1)
        if mask is not None:
            logits = logits + mask[:, None]
        p = self.attn_drop(logits.softmax(-1))
        return self.proj_drop(self.out_proj((p @ split(v)).transpose(1, 2).reshape(nWin, N, C)))


class WindowMSA(_WindowAttnBase):
    def __init__(self, dim, heads, window, attn_drop=0., proj_drop=0., with_rpe=True):
        super().__init__()
        self._init_common(dim, heads, window, with_rpe)
        self.qkv = nn.Linear(dim, 3 * dim)
        self.attn_drop = nn.Dropout(attn_drop)
        self.out_proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x, kv=None, mask=None):
        q, k, v = self.qkv(x).chunk(3, dim=-1)
        return self._core(q, k, v, mask)


class WindowMCA(_WindowAttnBase):
    def __init__(self, dim, heads, window, attn_drop=0., proj_drop=0., with_rpe=True):
        super().__init__()
        self._init_common(dim, heads, window, with_rpe)
        self.k_proj = nn.Linear(dim, dim)
        self.v_proj = nn.Linear(dim, dim)
        self.q_proj = nn.Linear(dim, dim)
        self.attn_drop = nn.Dropout(attn_drop)
        self.out_proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x, kv, mask=None):
        return self._core(self.q_proj(x), self.k_proj(kv), self.v_proj(kv), mask)


class WindowedAttention(nn.Module):
    """Pad -> partition -> window attention -> merge -> crop, for self-attention
    (`LocalWindowSelfAttention`) and cross-attention (`MultiWindowCrossAttention`).
    The inner module is registered as `.attn` like in the reference."""

    def __init__(self, dim, heads, window, cross, with_pad_mask=False, **kw):
        super().__init__()
        window = (window, window) if isinstance(window, int) else tuple(window)
        self.window_size = window
        self.with_pad_mask = with_pad_mask
        self.cross = cross
        self.attn = (WindowMCA if cross else WindowMSA)(dim, heads, window, **kw)

    def forward(self, x, kv, H, W):
        B, N, C = x.shape
        g = window_geometry(H, W, *self.window_size)
        Wh, Ww = self.window_size

        def part(t):
            t = F.pad(t.view(B, H, W, C), (0, 0, g.pad_l, g.pad_r, g.pad_t, g.pad_b))
            return (t.view(B, g.nWh, Wh, g.nWw, Ww, C).transpose(2, 3)
                    .reshape(-1, Wh * Ww, C))
        mask = None
        if self.with_pad_mask and g.pad_h > 0 and g.pad_w > 0:
            m = F.pad(x.new_zeros(1, H, W, 1), (0, 0, g.pad_l, g.pad_r, g.pad_t, g.pad_b),
                      value=-math.inf)
            m = m.view(1, g.nWh, Wh, g.nWw, Ww, 1).transpose(2, 3).reshape(-1, Wh * Ww)
            mask = m[:, None, :].expand(-1, Wh * Ww, -1).repeat(B, 1, 1)
        out = self.attn(part(x), part(kv) if self.cross else None, mask)
        out = out.view(B, g.nWh, g.nWw, Wh, Ww, C).transpose(2, 3).reshape(B, g.Hp, g.Wp, C)
        return out[:, g.pad_t:g.pad_t + H, g.pad_l:g.pad_l + W].reshape(B, N, C)


class CrossFFN(nn.Module):
    """1x1 -> BN -> GELU -> dw3x3 -> BN -> GELU -> 1x1 -> BN -> GELU (`.layers.0..8`)."""

    def __init__(self, cin, hidden, cout, norm_cfg):
        super().__init__()
        self.layers = NormActSequential(
            nn.Conv2d(cin, hidden, 1), make_norm(norm_cfg, hidden), nn.GELU(),
            make_conv(hidden, hidden, 3, 1, 1, groups=hidden), make_norm(norm_cfg, hidden),
            nn.GELU(),
            nn.Conv2d(hidden, cout, 1), make_norm(norm_cfg, cout), nn.GELU())

    def forward(self, x, H, W):
        return _tokens(self.layers(_image(x, H, W)))


class HRFormerBlock(nn.Module):
    expansion = 1

    def __init__(self, cin, cout, num_heads, window_size=7, mlp_ratio=4, drop_path=0.,
                 norm_cfg=None, transformer_norm_cfg=None, with_rpe=True,
                 with_pad_mask=False, **_ignored):
        super().__init__()
        self.num_heads, self.window_size, self.mlp_ratio = num_heads, window_size, mlp_ratio
        self.norm1 = make_norm(transformer_norm_cfg, cin)
        self.attn = WindowedAttention(cin, num_heads, window_size, cross=False,
                                      with_rpe=with_rpe, with_pad_mask=with_pad_mask)
        self.norm2 = make_norm(transformer_norm_cfg, cout)
        self.ffn = CrossFFN(cin, int(cin * mlp_ratio), cout, norm_cfg)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()

    def forward_tokens(self, t, H, W):
        """the block on (B, H*W, C) tokens: consecutive blocks of a branch stay in this layout"""
        t = t + self.drop_path(self.attn(self.norm1(t), None, H, W))
        return t + self.drop_path(self.ffn(self.norm2(t), H, W))

    def forward(self, x):
        B, C, H, W = x.shape
        return _image(self.forward_tokens(_tokens(x), H, W), H, W).contiguous()


def run_blocks(blocks, x):
    """nn.Sequential of HRFormerBlocks on an NCHW map with ONE layout round trip for the whole
    chain (the reference converts NCHW <-> tokens in every block, hrformer.py:365-373)"""
    if not all(isinstance(b, HRFormerBlock) for b in blocks):
        return blocks(x)
    B, C, H, W = x.shape
    t = _tokens(x)
    for b in blocks:
        t = b.forward_tokens(t, H, W)
    return _image(t, H, W).contiguous()


class HRFuserFusionBlock(nn.Module):
    """Camera tokens query each extra modality through its own MWCA; see
    reference hrfuser_hrformer_based.py:305-317 for the residual wiring."""
    expansion = 1

    def __init__(self, cin, cout, num_heads, window_size=7, mlp_ratio=4, drop_path=0.,
                 norm_cfg=None, transformer_norm_cfg=None, num_fused_modalities=2,
                 proj_drop_rate=0., with_cp=False, **_ignored):
        super().__init__()
        self.num_heads, self.window_size, self.mlp_ratio = num_heads, window_size, mlp_ratio
        self.num_fused_modalities = num_fused_modalities
        self.with_cp = with_cp
        M = num_fused_modalities
        self.norm1 = nn.ModuleList(make_norm(transformer_norm_cfg, cin) for _ in range(M))
        self.norm2 = nn.ModuleList(make_norm(transformer_norm_cfg, cout) for _ in range(M))
        self.attn = nn.ModuleList(
            WindowedAttention(cin, num_heads, window_size, cross=True, proj_drop=proj_drop_rate)
            for _ in range(M))
        self.norm3 = make_norm(transformer_norm_cfg, cout)
        self.ffn = CrossFFN(cin, int(cin * mlp_ratio), cout, norm_cfg)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()

    def forward(self, x, mods):
        if self.with_cp and x.requires_grad:
            raise Exception('with_cp is currently not possible with CA Fusion module')
        B, C, H, W = x.shape
        cam = _tokens(x)
        acc = cam
        for k in range(self.num_fused_modalities):
            z = _tokens(mods[k])
            acc = acc + z + self.drop_path(self.attn[k](self.norm1[k](cam), self.norm2[k](z), H, W))
        acc = acc + self.drop_path(self.ffn(self.norm3(acc), H, W))
        return _image(acc, H, W).contiguous()


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, norm_cfg, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = make_norm(norm_cfg, planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = make_norm(norm_cfg, planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = make_norm(norm_cfg, planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        y = norm_act(self.bn1, self.relu, self.conv1(x))
        y = norm_act(self.bn2, self.relu, self.conv2(y))
        y = self.bn3(self.conv3(y))
        return self.relu(y + (x if self.downsample is None else self.downsample(x)))


def make_bottleneck_layer(inplanes, planes, blocks, norm_cfg):
    down = None
    if inplanes != planes * 4:
        down = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, bias=False),
                             make_norm(norm_cfg, planes * 4))
    layers = [Bottleneck(inplanes, planes, norm_cfg, down)]
    layers += [Bottleneck(planes * 4, planes, norm_cfg) for _ in range(1, blocks)]
    return nn.Sequential(*layers)


class HRFormerModule(nn.Module):
    """Parallel branches of HRFormerBlocks followed by the multi-resolution
    exchange (reference hrnet.py:184-207 driving hrformer.py:498-561)."""

    def __init__(self, num_branches, num_blocks, channels, heads, windows, mlp_ratios,
                 multiscale_output, norm_cfg, transformer_norm_cfg, drop_paths,
                 with_rpe=True, with_pad_mask=False):
        super().__init__()
        if not (num_branches == len(num_blocks) == len(channels)):
            raise ValueError(f'NUM_BRANCHES({num_branches}) does not match '
                             f'NUM_BLOCKS({len(num_blocks)}) / NUM_CHANNELS({len(channels)})')
        self.num_branches = num_branches
        self.in_channels = list(channels)
        self.multiscale_output = multiscale_output
        self.branches = nn.ModuleList(
            nn.Sequential(*[
                HRFormerBlock(channels[b], channels[b], heads[b], windows[b], mlp_ratios[b],
                              drop_path=drop_paths[i], norm_cfg=norm_cfg,
                              transformer_norm_cfg=transformer_norm_cfg, with_rpe=with_rpe,
                              with_pad_mask=with_pad_mask)
                for i in range(num_blocks[b])])
            for b in range(num_branches))
        self.fuse_layers = self._exchange_layers(norm_cfg)
        self.relu = nn.ReLU(inplace=False)

    def _exchange_layers(self, norm_cfg):
        nb, ch = self.num_branches, self.in_channels
        if nb == 1:
            return None
        rows = []
        for i in range(nb if self.multiscale_output else 1):
            row = []
            for j in range(nb):
                if j == i:
                    row.append(None)
                elif j > i:       # coarser -> finer: 1x1 + BN, upsampled in forward
                    row.append(conv_bn(ch[j], ch[i], 1, 1, norm_cfg, relu=False))
                else:             # finer -> coarser: (i-j) x [dw3x3 s2, BN, 1x1, BN, (ReLU)]
                    chain = []
                    for t in range(i - j):
                        last = t == i - j - 1
                        cout = ch[i] if last else ch[j]
                        step = [make_conv(ch[j], ch[j], 3, 2, 1, groups=ch[j], bias=False),
                                make_norm(norm_cfg, ch[j]),
                                nn.Conv2d(ch[j], cout, 1, bias=False),
                                make_norm(norm_cfg, cout)]
                        if not last:
                            step.append(nn.ReLU(False))
                        chain.append(NormActSequential(*step))
                    row.append(nn.Sequential(*chain))
            rows.append(nn.ModuleList(row))
        return nn.ModuleList(rows)

    def forward(self, xs):
        if self.num_branches == 1:
            return [run_blocks(self.branches[0], xs[0])]
        xs = [run_blocks(br, x) for br, x in zip(self.branches, xs)]
        outs = []
        for i, row in enumerate(self.fuse_layers):
            acc = xs[i]
            for j in range(self.num_branches):
                if j > i:
                    acc = acc + F.interpolate(row[j](xs[j]), size=xs[i].shape[2:],
                                              mode='bilinear', align_corners=False)
                elif j < i:
                    acc = acc + row[j](xs[j])
            outs.append(self.relu(acc))
        return outs


def make_transition(pre, cur, norm_cfg):
    """reference hrnet.py:419-463"""
    layers = []
    for i, c in enumerate(cur):
        if i < len(pre):
            layers.append(None if c == pre[i] else conv_bn(pre[i], c, 3, 1, norm_cfg, 'inplace'))
        else:
            steps = i + 1 - len(pre)
            layers.append(nn.Sequential(*[
                conv_bn(pre[-1], c if s == steps - 1 else pre[-1], 3, 2, norm_cfg, 'inplace')
                for s in range(steps)]))
    return nn.ModuleList(layers)
