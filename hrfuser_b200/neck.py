"""HRFPN neck, drop-in for `mmdet/models/necks/hrfpn.py:13-100`.

Same class / registry name (`HRFPN`), ctor arguments and state_dict layout
(`reduction_conv.conv.{weight,bias}`, `fpn_convs.{i}.conv.{weight,bias}`: mmcv `ConvModule`
without norm / activation), same `forward(inputs) -> tuple of num_outs maps`.

The reference's front half (`hrfpn.py:77-86`) upsamples the three coarse maps to the finest
grid, concatenates all four (270 channels at 96 x 160 for HRFuser-T: 133 MB per 8 frames in
fp32, the largest tensor after the backbone) and runs the 1x1 `reduction_conv` on it.
Bilinear interpolation is linear and acts per channel, so it commutes with a 1x1 convolution:

    W . cat(x0, up(x1), up(x2), up(x3)) + b  ==  (W0 x0 + b) + up(W1 x1) + up(W2 x2) + up(W3 x3)

with W_i the column slice of W for branch i.  Eval forward on the GPU therefore runs four 1x1
convolutions at their *own* resolution (`hrf_pw_fwd`: 15x fewer FLOPs, 2.1 instead of 17 GFLOP
per 8 frames) and one `hrf_fuse_sum_fwd` (the exchange kernel of the backbone: x + bilinear
gathers, here without ReLU); the upsampled maps and the concatenation never exist.  The
pooling pyramid and the 3x3 `fpn_convs` (dense 256 -> 256 convolutions: cuDNN territory, like
the Bottlenecks of SURVEY 8 a11) stay torch ops on the result.

No fallback: the eval forward raises without the built library / a CUDA tensor
(`_lib.HrfError`); `train()` / grad-enabled calls take the reference's formulation in torch
autograd, as the backbone does.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops


class _ConvModule(nn.Module):
    """mmcv `ConvModule(..., norm_cfg=None, act_cfg=None)`: a conv under the attribute `conv`."""

    def __init__(self, cin, cout, k, padding=0, stride=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=padding)

    def forward(self, x):
        return self.conv(x)


class HRFPN(nn.Module):
    def __init__(self, in_channels, out_channels, num_outs=5, pooling_type='AVG', conv_cfg=None,
                 norm_cfg=None, with_cp=False, stride=1,
                 init_cfg=dict(type='Caffe2Xavier', layer='Conv2d'), precision='fp32'):
        super().__init__()
        assert isinstance(in_channels, list)                     # hrfpn.py:49
        if conv_cfg is not None or norm_cfg is not None:
            raise NotImplementedError('HRFPN: conv_cfg / norm_cfg other than None (no shipped '
                                      'hrfuser config sets them)')
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.num_ins = len(in_channels)
        self.num_outs = num_outs
        self.with_cp = with_cp
        self.conv_cfg, self.norm_cfg, self.init_cfg = conv_cfg, norm_cfg, init_cfg
        self.precision = precision
        self.reduction_conv = _ConvModule(sum(in_channels), out_channels, 1)
        self.fpn_convs = nn.ModuleList(
            _ConvModule(out_channels, out_channels, 3, padding=1, stride=stride)
            for _ in range(num_outs))
        self.pooling = F.max_pool2d if pooling_type == 'MAX' else F.avg_pool2d
        self._blobs, self._blobs_ver = None, None
        self._fpn_blobs, self._fpn_ver = None, None

    def init_weights(self):
        """Caffe2Xavier on every Conv2d (mmcv: kaiming_uniform, a=1, fan_in, bias 0)."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, a=1, mode='fan_in', nonlinearity='leaky_relu')
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    # ------------------------------------------------------------ packed weights
    def invalidate(self):
        self._blobs = None
        self._fpn_blobs = None

    def train(self, mode=True):
        # an optimizer updates reduction_conv in place between two evaluations: never carry
        # packed weights across a train() phase (the backbone does the same with its engine)
        self._blobs = self._fpn_blobs = None
        return super().train(mode)

    def _load_from_state_dict(self, *a, **k):
        # reached for every module of the tree when a PARENT (the detector, mmcv's
        # load_checkpoint) loads a checkpoint; Module.load_state_dict of a child is not
        self._blobs = self._fpn_blobs = None
        return super()._load_from_state_dict(*a, **k)

    def _weights_version(self):
        conv = self.reduction_conv.conv
        return (conv.weight._version, conv.weight.data_ptr(),
                None if conv.bias is None else (conv.bias._version, conv.bias.data_ptr()))

    def _apply(self, fn, *a, **k):
        self._blobs = self._fpn_blobs = None
        return super()._apply(fn, *a, **k)

    def _packed(self, device):
        """per-branch column slices of the reduction conv as `hrf_pw` blobs (bias on branch 0)"""
        # keyed on the parameters' version counters as well: any in-place update repacks
        ver = self._weights_version()
        if self._blobs is None or self._blobs[0].device != device or self._blobs_ver != ver:
            conv = self.reduction_conv.conv
            blobs, off = [], 0
            for i, c in enumerate(self.in_channels):
                part = nn.Conv2d(c, self.out_channels, 1, bias=(i == 0 and conv.bias is not None))
                with torch.no_grad():
                    part.weight.copy_(conv.weight[:, off:off + c])
                    if part.bias is not None:
                        part.bias.copy_(conv.bias)
                blobs.append(ops.pack_pw(part, None).to(device))
                off += c
            self._blobs, self._blobs_ver = blobs, ver
        return self._blobs

    # ------------------------------------------------------------------ forward
    def reduce(self, inputs, as_tokens=False):
        """`reduction_conv(cat(x0, up(x1), ...))` of hrfpn.py:79-86 -> fp32 NCHW (or, with
        `as_tokens`, the kernels' own [B,H,W,C] tokens in the working precision)"""
        x0 = inputs[0]
        if not x0.is_cuda:
            raise _lib.HrfError('HRFPN: the eval forward runs on the hrfuser_b200 CUDA kernels only '
                                '(no CPU fallback); move the inputs to a CUDA device')
        for i, t in enumerate(inputs[1:], 1):          # F.interpolate(scale_factor=2**i)
            if (t.shape[2] << i, t.shape[3] << i) != tuple(x0.shape[2:]):
                raise ValueError(f'HRFPN: input {i} is {tuple(t.shape[2:])}, expected '
                                 f'{(x0.shape[2] >> i, x0.shape[3] >> i)}')
        dt = torch.bfloat16 if self.precision == 'bf16' else torch.float32
        blobs = self._packed(x0.device)
        # one stream: forking the three coarse (latency-bound) branches onto side streams was
        # measured and dropped (profiles/r01_neck_bench_streams.jsonl: HRFuser-B 0.85 -> 0.71 ms,
        # HRFuser-T 0.41 -> 0.39 ms in bf16 but 0.69 -> 0.87 ms in fp32)
        ys = [ops.pointwise(ops.nchw_to_nhwc(t.contiguous(), dtype=dt), blob, self.out_channels)
              for t, blob in zip(inputs, blobs)]
        out = ops.fuse_sum(ys[0], ups=ys[1:], relu=False) if len(ys) > 1 else ys[0]
        return out if as_tokens else ops.nhwc_to_nchw(out, dtype=torch.float32)

    def _back_half_blobs(self, device):
        """fpn_convs as `hrf_convgemm` blobs (bf16 mode), keyed like the reduction slices"""
        ver = tuple((c.conv.weight._version, c.conv.weight.data_ptr(),
                     None if c.conv.bias is None else c.conv.bias._version) for c in self.fpn_convs)
        if self._fpn_blobs is None or self._fpn_blobs[0].device != device or self._fpn_ver != ver:
            self._fpn_blobs = [ops.pack_convgemm(c.conv, None).to(device) for c in self.fpn_convs]
            self._fpn_ver = ver
        return self._fpn_blobs

    def _tc_back_half(self, x0):
        """True when pooling + the 3x3 convs run on the library's kernels: bf16 mode, channel
        count covered by hrf_convgemm_fwd, map divisible by the pyramid's windows."""
        oc = self.out_channels
        return (self.precision == 'bf16' and all(c.conv.stride == (1, 1) for c in self.fpn_convs) and
                ops.convgemm_supported(oc, oc, 3, 1) and oc % 8 == 0 and
                x0.shape[2] % (1 << (self.num_outs - 1)) == 0 and x0.shape[3] % (1 << (self.num_outs - 1)) == 0 and
                not any(getattr(c, 'activate', None) is not None for c in self.fpn_convs))

    def forward(self, inputs):
        assert len(inputs) == self.num_ins                      # hrfpn.py:78
        if self.training or torch.is_grad_enabled() and (
                any(t.requires_grad for t in inputs) or any(p.requires_grad for p in self.parameters())):
            # (eval() with trainable parameters and grad enabled -- fine-tuning with frozen BN
            # statistics -- must build a graph, as the reference does)
            return self._forward_autograd(inputs)
        # fp32 mode: the cuDNN 3x3 convs must not drop to TF32 (3e-4 off; the engine does the same)
        if inputs[0].is_cuda and self._tc_back_half(inputs[0]):
            # bf16 mode: the whole neck on the library's kernels -- the reduced map stays in bf16
            # tokens, the pyramid levels come from hrf_pool_fwd, the five 3x3 convs from the
            # TMA / tcgen05 implicit-GEMM kernel (hrf_convgemm_fwd); fp32 NCHW only at the exit
            with torch.no_grad():
                tok = self.reduce(list(inputs), as_tokens=True)
                blobs = self._back_half_blobs(tok.device)
                mode = 'MAX' if self.pooling is F.max_pool2d else 'AVG'
                levels = [tok] + [ops.pool_tokens(tok, 2 ** i, mode) for i in range(1, self.num_outs)]
                return tuple(ops.nhwc_to_nchw(ops.conv_gemm(lv, blobs[i], self.out_channels, 3, 1, False),
                                              dtype=torch.float32) for i, lv in enumerate(levels))
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=True,
                                                         allow_tf32=self.precision != 'fp32'):
            out = self.reduce(list(inputs))
            outs = [out] + [self.pooling(out, kernel_size=2 ** i, stride=2 ** i)
                            for i in range(1, self.num_outs)]
            return tuple(self.fpn_convs[i](outs[i]) for i in range(self.num_outs))

    def _forward_autograd(self, inputs):
        """training path: the reference's formulation (hrfpn.py:77-100) on torch autograd"""
        outs = [inputs[0]]
        for i in range(1, self.num_ins):
            outs.append(F.interpolate(inputs[i], scale_factor=2 ** i, mode='bilinear'))
        out = self.reduction_conv(torch.cat(outs, dim=1))
        outs = [out] + [self.pooling(out, kernel_size=2 ** i, stride=2 ** i)
                        for i in range(1, self.num_outs)]
        return tuple(self.fpn_convs[i](outs[i]) for i in range(self.num_outs))


def register_with_mmdet():
    """Register the class under its reference name when mmdet is importable."""
    try:
        from mmdet.models.builder import NECKS
    except Exception:
        return False
    NECKS.register_module(name='HRFPN', force=True, module=HRFPN)
    return True
