"""Inference engine: walks a `HRFuserHRFormerBased` module once, packs its
parameters for the sm_100a kernels, and runs the forward with them.

What runs where:
  * window attention (LSA / MWCA), MixFFN, the multi-resolution exchange and the
    pad/partition/merge reshapes  -> hand-written kernels behind the C-ABI (`ops`)
  * stems, Bottlenecks and 3x3 transition convs ("next")
    -> cuDNN through torch, BatchNorm folded into the conv, channels-last
Activations between the two worlds are channels-last, so no layout copy is made:
a (B,C,H,W) channels_last tensor *is* the (B,H,W,C) token tensor the kernels read.

The wiring follows the reference forward (hrfuser_hrformer_based.py:522-627),
including the `transition1[i][0]` quirk (:550-551).
"""
import contextlib
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops


def _fold_conv_bn(conv, bn, dtype):
    """conv (no bias or bias) followed by eval-mode BN -> (weight, bias) tensors."""
    w = conv.weight.detach().float()
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
    if bn is not None:
        s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
        w = w * s.view(-1, 1, 1, 1)
        b = (b - bn.running_mean.detach().float()) * s + bn.bias.detach().float()
    return (w.to(dtype).contiguous(memory_format=torch.channels_last), b.to(dtype))


class _Conv:
    """Folded conv+BN(+ReLU)(+residual) executed by cuDNN, channels-last.  On CUDA the
    bias / residual / ReLU epilogue is fused into the convolution
    (`cudnn_convolution_relu`, `cudnn_convolution_add_relu`): torch's separate
    broadcast-add and clamp passes cost 3-4x the convolution itself on the
    256-channel stem tensors (profiles/r01_v1_simt_launches.txt)."""

    def __init__(self, conv, bn, relu, dtype, extra_bias=None, use_bias=True, engine=None):
        # bf16 mode on the GPU: the library's warp-specialised TMA / tcgen05 implicit-GEMM kernel
        # (csrc/conv_gemm_tc.cuh) when it covers the layer -- every stem / Bottleneck / 256-channel
        # transition conv of the shipped configs; HRF_CONVGEMM=0 keeps cuDNN (A/B runs)
        k = conv.kernel_size[0]
        self.tc = (engine is not None and engine.ops is ops and dtype == torch.bfloat16 and
                   conv.weight.is_cuda and os.environ.get('HRF_CONVGEMM', '1') != '0' and
                   conv.groups == 1 and conv.kernel_size == (k, k) and conv.padding == (k // 2, k // 2) and
                   conv.dilation == (1, 1) and conv.stride[0] == conv.stride[1] and
                   ops.convgemm_supported(conv.in_channels, conv.out_channels, k, conv.stride[0]))
        self.relu, self.cout = relu, conv.out_channels
        if self.tc:
            blob = ops.pack_convgemm(conv, bn, bn.eps if bn is not None else 1e-5, extra_bias)
            if not use_bias:                        # bias-free branch: its shift rides on another conv's bias
                npad = 32 if self.cout <= 32 else 64 if self.cout <= 64 else 128 if self.cout <= 128 else 256
                blob[:npad] = 0
            self.blob, self.ksize, self.stride1 = engine._blob(blob), k, conv.stride[0]
            return
        self.w, b = _fold_conv_bn(conv, bn, dtype)
        if extra_bias is not None:
            b = b + extra_bias
        self.has_bias = use_bias and (conv.bias is not None or bn is not None or extra_bias is not None)
        self.stride, self.padding, self.groups, self.relu = conv.stride, conv.padding, conv.groups, relu
        # cuDNN has no good kernel for output widths like 18 / 36 (it falls back to slower
        # engines plus nhwcAddPadding passes: 65.6 -> 43.5 us and 33.7 -> 17.7 us for the
        # 256 -> 18 / 36 transitions, tools/probe_trans1.py): compute zero-padded output
        # channels and slice them off afterwards
        self.cout = self.w.shape[0]
        if self.w.is_cuda and self.cout % 8 != 0 and self.groups == 1:
            padded = (self.cout + 15) // 16 * 16
            w = self.w.new_zeros((padded,) + tuple(self.w.shape[1:]))
            w[:self.cout] = self.w
            self.w = w.contiguous(memory_format=torch.channels_last)
            b = torch.cat([b, b.new_zeros(padded - self.cout)])
        self.b = b if self.has_bias else None
        self.b32 = b.float() if self.has_bias else None

    def __call__(self, x, residual=None):
        if self.tc:
            t = x.permute(0, 2, 3, 1)
            r = residual.permute(0, 2, 3, 1) if residual is not None else None
            y = ops.conv_gemm(t if t.is_contiguous() else t.contiguous(), self.blob.t, self.cout, self.ksize,
                              self.stride1, self.relu, r if r is None or r.is_contiguous() else r.contiguous())
            return y.permute(0, 3, 1, 2)
        y = self._run(x, residual)
        if y.shape[1] != self.cout:
            y = y[:, :self.cout].contiguous(memory_format=torch.channels_last)
        return y

    def _run(self, x, residual=None):
        if residual is not None and residual.shape[1] != self.w.shape[0]:
            # output channels were zero-padded (widths cuDNN has no kernel for): pad the residual alike
            residual = F.pad(residual, (0, 0, 0, 0, 0, self.w.shape[0] - residual.shape[1])).contiguous(
                memory_format=torch.channels_last)
        if x.is_cuda and self.relu and self.has_bias:
            if residual is None:
                return torch.cudnn_convolution_relu(x, self.w, self.b, self.stride, self.padding,
                                                    (1, 1), self.groups)
            return torch.cudnn_convolution_add_relu(x, self.w, residual, 1.0, self.b, self.stride,
                                                    self.padding, (1, 1), self.groups)
        if x.is_cuda and (self.has_bias or residual is not None or self.relu):
            y = F.conv2d(x, self.w, None, self.stride, self.padding, 1, self.groups)
            t = y.permute(0, 2, 3, 1)
            if t.is_contiguous():
                r = residual.permute(0, 2, 3, 1) if residual is not None else None
                bias = self.b32 if self.has_bias else torch.zeros(y.shape[1], device=y.device)
                ops.bias_act_(t, bias, r if r is None or r.is_contiguous() else r.contiguous(),
                              self.relu)
                return y
        y = F.conv2d(x, self.w, self.b, self.stride, self.padding, 1, self.groups)
        if residual is not None:
            y += residual
        return y.relu_() if self.relu else y


def _seq(mods, dtype, engine=None):
    """nn.Sequential of [conv, bn, (relu)] -> _Conv"""
    mods = list(mods)
    relu = len(mods) > 2 and isinstance(mods[2], nn.ReLU)
    return _Conv(mods[0], mods[1], relu, dtype, engine=engine)


class _Bottleneck:
    """conv1-bn-relu, conv2-bn-relu, conv3-bn (+ downsample conv-bn) + add + relu
    (reference resnet.py:263-302).  The downsample branch runs bias-free and its
    folded BN shift moves into conv3's bias, so the add + ReLU fuse into conv3."""

    def __init__(self, m, dtype, engine=None):
        self.c1 = _Conv(m.conv1, m.bn1, True, dtype, engine=engine)
        self.c2 = _Conv(m.conv2, m.bn2, True, dtype, engine=engine)
        self.down, extra = None, None
        if m.downsample is not None:
            _, extra = _fold_conv_bn(m.downsample[0], m.downsample[1], torch.float32)
            self.down = _Conv(m.downsample[0], m.downsample[1], False, dtype, use_bias=False, engine=engine)
        # ReLU after the add
        self.c3 = _Conv(m.conv3, m.bn3, True, dtype, extra_bias=extra, engine=engine)
        if not self.c3.tc and extra is not None:
            self.c3 = _Conv(m.conv3, m.bn3, True, dtype, extra_bias=extra.to(dtype))
        # conv3 and the downsample conv are both 1x1 convs into the same sum: ONE GEMM over the
        # concatenated K = planes + inplanes (hrf_convgemm_grouped_cat_fwd); the downsample map
        # (63 MB per stream at 96 x 160 x 8) is never written or re-read.  HRF_CONV_CAT=0: two launches
        self.cat = None
        d = m.downsample[0] if m.downsample is not None else None
        if (d is not None and self.c3.tc and self.down.tc and d.kernel_size == (1, 1) and d.stride == (1, 1) and
                m.conv3.in_channels % 64 == 0 and d.in_channels % 64 == 0 and
                os.environ.get('HRF_CONV_CAT', '1') != '0'):
            self.cat = _Conv(fold_conv3_downsample(m), None, True, dtype, engine=engine)
            self.cat.cin1 = m.conv3.in_channels

    def __call__(self, x):
        if self.cat is not None:
            return run_cat([self.cat], [self.c2(self.c1(x))], [x])[0]
        idt = x if self.down is None else self.down(x)
        return self.c3(self.c2(self.c1(x)), residual=idt)


def fold_conv3_downsample(m):
    """Bottleneck `bn3(conv3(y)) + bn_d(downsample(x))` (reference resnet.py:287-300, eval-mode BN) as
    ONE 1x1 convolution of the channel concatenation [y | x]: weights [W3 * s3 | Wd * sd], bias
    b3 + bd."""
    d, bnd = m.downsample[0], m.downsample[1]
    w3, b3 = _fold_conv_bn(m.conv3, m.bn3, torch.float32)
    wd, bd = _fold_conv_bn(d, bnd, torch.float32)
    both = nn.Conv2d(m.conv3.in_channels + d.in_channels, m.conv3.out_channels, 1, bias=True)
    with torch.no_grad():
        both.weight.copy_(torch.cat([w3, wd], 1))
        both.bias.copy_(b3 + bd)
    return both.to(m.conv3.weight.device)


def run_cat(convs, ys, xs):
    """relu(conv3(y) + downsample(x)) of one layer of n streams as one concatenated-K GEMM launch"""
    tok = BackboneEngine._tokens
    outs = ops.conv_gemm_cat_grouped([tok(y) for y in ys], [tok(x) for x in xs], [c.blob.t for c in convs],
                                     convs[0].cout, [c.relu for c in convs])
    return [BackboneEngine._image(o) for o in outs]


class BackboneEngine:
    def __init__(self, module, precision='fp32', device_ops=None):
        """`device_ops` replaces the forward ops of `ops` (the packers always come
        from the real library); only the CPU test-suite passes it, to check the
        packing and the wiring against a blob-level emulation without a GPU."""
        if precision not in ('fp32', 'bf16'):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        lib = _lib.load()                       # raises if the CUDA library is not built
        dev = module.conv1.weight.device
        self.ops = device_ops or ops
        if device_ops is None:
            if dev.type != 'cuda':
                raise _lib.HrfError('HRFuserHRFormerBased inference needs a CUDA device: the '
                                    'eval forward runs on sm_100a kernels and has no CPU '
                                    'fallback (call .cuda() first)')
            with torch.cuda.device(dev):
                _lib.check(lib.hrf_device_check())
        if any(isinstance(m, nn.modules.batchnorm._BatchNorm) and m.training
               for m in module.modules()):
            raise _lib.HrfError('engine built while BatchNorm layers are in training mode')
        self.device, self.precision = dev, precision
        self.dtype = torch.float32 if precision == 'fp32' else torch.bfloat16
        self.M = module.num_fused_modalities
        self.pad_mask = module.with_pad_mask
        self._host_blobs, self._blob_slots = [], []
        m, dt = module, self.dtype

        self.stem = [self._stem_conv(m.conv1, m.bn1), _Conv(m.conv2, m.bn2, True, dt, engine=self)] + \
                    [_Bottleneck(b, dt, self) for b in m.layer1]
        self.stem_mod = [[self._stem_conv(m.conv_a[k], m.norm_a[k]),
                          _Conv(m.conv_b[k], m.norm_b[k], True, dt, engine=self)] +
                         [_Bottleneck(b, dt, self) for b in m.layer_a[k]] for k in range(self.M)]
        # transition1[i][0]: bare conv on branch 0, full conv-bn-relu on branch 1
        self.trans1 = [_Conv(m.transition1[0][0], None, False, dt, engine=self)] + \
                      [_seq(m.transition1[i][0], dt, self) for i in range(1, len(m.transition1))]
        self.trans_cam = {2: self._transitions(m.transition2), 3: self._transitions(m.transition3)}
        letters = ['a', 'b', 'c'] + (['d'] if m.pre_neck_fusion else [])
        self.trans_mod = {l: [self._transitions(getattr(m, f'transition_{l}')[k])
                              for k in range(self.M)] for l in letters}
        self.fusion = {l: [self._fusion_block(b) for b in getattr(m, f'fusion_{l}')]
                       for l in letters}
        self.stage = {i: self._stage(getattr(m, f'stage{i}')) for i in (2, 3, 4)}
        self.stage_mod = {l: [self._stage(getattr(m, f'stage_{l}')[k]) for k in range(self.M)]
                          for l in letters[1:]}
        self.pre_neck_fusion = m.pre_neck_fusion
        # independent branches / streams run on forked CUDA streams (see _par)
        self.concurrent = device_ops is None and os.environ.get('HRF_SERIAL', '0') != '1'
        # frames per sub-batch of the stem phase (0 = the whole batch at once)
        self.stem_chunk = int(os.environ.get('HRF_STEM_CHUNK', '0'))
        self._free_streams, self._keep = [], []
        self._upload()

    # ---------------------------------------------------------------- packing
    def _stem_conv(self, conv, bn):
        """First conv of a stream.  bf16 mode: the tcgen05 stem kernel, fed with the
        caller's fp32 NCHW image (no cast / layout passes); otherwise cuDNN."""
        tc = (self.ops is ops and self.precision == 'bf16' and conv.in_channels <= 3 and
              conv.out_channels == 64 and conv.kernel_size == (3, 3) and conv.stride == (2, 2) and
              conv.padding == (1, 1) and conv.bias is None and conv.groups == 1)
        if not tc:
            cv = _Conv(conv, bn, True, self.dtype)
            return lambda x: cv(self._prep(x))
        blob = self._blob(ops.pack_stem(conv, bn, bn.eps))
        return lambda x: self._image(ops.stem_conv(
            x.to(device=self.device, dtype=torch.float32).contiguous(), blob.t, 64, relu=True))

    def _blob(self, host_blob):
        """register a host blob; returns a slot whose .t is the device view after upload"""
        slot = type('Blob', (), {})()
        self._host_blobs.append(host_blob)
        self._blob_slots.append(slot)
        return slot

    def _upload(self):
        """one H2D copy for all packed weights; 256-byte aligned sub-views"""
        offs, total = [], 0
        for b in self._host_blobs:
            offs.append(total)
            total += (b.numel() + 63) // 64 * 64
        arena = torch.zeros(max(total, 64), dtype=torch.float32)
        for b, o in zip(self._host_blobs, offs):
            arena[o:o + b.numel()] = b
        self.arena = arena.to(self.device)
        for slot, b, o in zip(self._blob_slots, self._host_blobs, offs):
            slot.t = self.arena[o:o + b.numel()]
        self._host_blobs = None

    def _tconv(self, mods):
        """[conv3x3, bn, relu] of a transition layer.  Narrow ones (the modality transitions
        18 -> 36 -> 72 -> 144 and the camera's new-branch convs) run in the library's implicit
        GEMM kernel on tokens: cuDNN has no good kernel for these channel counts and wraps
        each call in ~10 padding / layout launches."""
        mods = list(mods)
        conv, bn = mods[0], mods[1]
        relu = len(mods) > 2 and isinstance(mods[2], nn.ReLU)
        own = (isinstance(bn, nn.BatchNorm2d) and conv.kernel_size == (3, 3) and conv.groups == 1 and
               conv.padding == (1, 1) and conv.stride in ((1, 1), (2, 2)) and conv.dilation == (1, 1) and
               conv.in_channels <= 72 and conv.in_channels % 2 == 0 and conv.out_channels % 2 == 0)
        if not own:
            return _Conv(conv, bn, relu, self.dtype, engine=self)
        blob = self._blob(ops.pack_conv3x3(conv, bn, bn.eps))
        cout, stride = conv.out_channels, conv.stride[0]
        return lambda x_img: self._image(self.ops.conv3x3(self._tokens(x_img), blob.t, cout, stride, relu))

    def _transitions(self, tl):
        out = []
        for t in tl:
            if t is None:
                out.append(None)
            elif isinstance(t[0], nn.Conv2d):
                out.append([self._tconv(t)])
            else:
                out.append([self._tconv(s) for s in t])
        return out

    def _ffn(self, ln, ffn):
        L = ffn.layers
        return dict(blob=self._blob(ops.pack_ffn(ln, L[0], L[1], L[3], L[4], L[6], L[7], L[1].eps)),
                    hidden=L[0].weight.shape[0], eps=ln.eps)

    def _hrformer_block(self, b):
        a = b.attn.attn
        Cc = a.qkv.weight.shape[1]
        wq, wk, wv = a.qkv.weight.detach().split(Cc, 0)
        bq, bk, bv = a.qkv.bias.detach().split(Cc, 0)
        ln = (b.norm1.weight, b.norm1.bias)
        blob = ops.pack_attn(Cc, a.num_heads, a.Wh, ln, ln, wq, bq, wk, bk, wv, bv,
                             a.out_proj.weight, a.out_proj.bias,
                             a.relative_position_bias_table if a.with_rpe else None)
        assert a.Wh == a.Ww
        return dict(attn=[self._blob(blob)], heads=a.num_heads, win=a.Wh, eps=b.norm1.eps,
                    pad_mask=b.attn.with_pad_mask, ffn=self._ffn(b.norm2, b.ffn))

    def _fusion_block(self, b):
        blobs = []
        for k in range(b.num_fused_modalities):
            a = b.attn[k].attn
            Cc = a.q_proj.weight.shape[0]
            blobs.append(self._blob(ops.pack_attn(
                Cc, a.num_heads, a.Wh, (b.norm1[k].weight, b.norm1[k].bias),
                (b.norm2[k].weight, b.norm2[k].bias), a.q_proj.weight, a.q_proj.bias,
                a.k_proj.weight, a.k_proj.bias, a.v_proj.weight, a.v_proj.bias,
                a.out_proj.weight, a.out_proj.bias,
                a.relative_position_bias_table if a.with_rpe else None)))
        a = b.attn[0].attn
        return dict(attn=blobs, heads=a.num_heads, win=a.Wh, eps=b.norm1[0].eps,
                    pad_mask=b.attn[0].with_pad_mask, ffn=self._ffn(b.norm3, b.ffn))

    def _stage(self, seq):
        mods = []
        for hm in seq:
            branches = [[self._hrformer_block(b) for b in br] for br in hm.branches]
            rows = None
            if hm.fuse_layers is not None:
                rows = []
                for i, row in enumerate(hm.fuse_layers):
                    ups, downs = [], []
                    for j, layer in enumerate(row):
                        if j > i:
                            ups.append((j, self._blob(ops.pack_pw(layer[0], layer[1], layer[1].eps)),
                                        layer[0].weight.shape[0]))
                        elif j < i:
                            chain = [(self._blob(ops.pack_dwpw(s[0], s[1], s[2], s[3], s[1].eps)),
                                      s[2].weight.shape[0], len(s) > 4) for s in layer]
                            downs.append((j, chain))
                    rows.append((ups, downs))
            mods.append((branches, rows))
        return mods

    # ---------------------------------------------------------------- running
    @staticmethod
    def _tokens(x):
        """(B,C,H,W) channels_last -> (B,H,W,C) contiguous view"""
        t = x.permute(0, 2, 3, 1)
        return t if t.is_contiguous() else t.contiguous()

    @staticmethod
    def _image(t):
        """(B,H,W,C) contiguous -> (B,C,H,W) channels_last view"""
        return t.permute(0, 3, 1, 2)

    # -- stream-level concurrency --------------------------------------------------
    # Branches of an HR module, the camera / modality streams and the rows of the
    # exchange step are independent; the low-resolution ones launch grids far
    # smaller than 148 SMs.  `_par` runs independent thunks on forked CUDA streams
    # (fork = side.wait_stream(cur), join = cur.wait_stream(side)); under CUDA-graph
    # capture the forks become parallel graph branches.  Results are kept alive
    # until the forward ends so the caching allocator never recycles a block that
    # another stream may still read.
    def _par(self, thunks):
        if not self.concurrent or len(thunks) <= 1:
            return [t() for t in thunks]
        cur = torch.cuda.current_stream()
        outs, used = [None] * len(thunks), []
        for idx in range(1, len(thunks)):
            side = self._free_streams.pop() if self._free_streams else torch.cuda.Stream(self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                outs[idx] = thunks[idx]()
            used.append(side)
        outs[0] = thunks[0]()
        for side in used:
            cur.wait_stream(side)
            self._free_streams.append(side)
        self._keep.append(outs)
        return outs

    def _run_block(self, blk, x, kv=None):
        y = self.ops.window_attention(x, kv, [s.t for s in blk['attn']], blk['heads'], blk['win'],
                                      blk['pad_mask'], blk['eps'])
        f = blk['ffn']
        return self.ops.mixffn(y, f['blob'].t, f['hidden'], f['eps'])

    def _run_branch(self, blocks, x):
        for blk in blocks:
            x = self._run_block(blk, x)
        return x

    def _exchange_row(self, i, ups, downs, ys, want_nchw):
        up_t = [self.ops.pointwise(ys[j], blob.t, cout) for j, blob, cout in ups]
        same_t = []
        for j, chain in downs:
            t = ys[j]
            for blob, cout, relu in chain:
                t = self.ops.dw_down(t, blob.t, cout, relu)
            same_t.append(t)
        return self.ops.fuse_sum(ys[i], up_t, same_t, relu=True, nchw_out=want_nchw)

    # -- the modality streams in lockstep with the camera's branch 0 -----------------------------
    # In stages 2 and 3 the camera's branch 0 and every modality stream walk the same number of
    # HRFormer blocks on tensors of one shape (each with its own weights).  Stepping through them
    # together turns (1 + M) attention and (1 + M) MixFFN launches per block into ONE of each
    # (hrf_window_attn_grouped_fwd / hrf_mixffn_grouped_fwd): launch, setup and the partly filled
    # last round of the persistent grids are paid once (3 x 640 MixFFN tiles over 296 resident
    # CTAs: 7 rounds instead of 3 x 3).  HRF_LOCKSTEP=0 runs the streams on their own.
    @staticmethod
    def _groupable(blocks, xs):
        b0, x0 = blocks[0], xs[0]
        key = lambda b: (b['heads'], b['win'], b['eps'], b['pad_mask'], b['ffn']['hidden'], b['ffn']['eps'], len(b['attn']))
        return (len(blocks) > 1 and x0.dtype == torch.bfloat16 and x0.shape[-1] == 18 and
                all(key(b) == key(b0) for b in blocks) and all(tuple(x.shape) == tuple(x0.shape) for x in xs))

    def _run_block_group(self, blocks, xs):
        if not self._groupable(blocks, xs):
            return [self._run_block(b, x) for b, x in zip(blocks, xs)]
        b0 = blocks[0]
        ys = self.ops.window_attention_grouped(xs, [b['attn'][0].t for b in blocks], b0['heads'], b0['win'],
                                               b0['pad_mask'], b0['eps'])
        return self.ops.mixffn_grouped(ys, [b['ffn']['blob'].t for b in blocks], b0['ffn']['hidden'], b0['ffn']['eps'])

    def _run_branch_lockstep(self, blocks, x, comp):
        """branch 0 of a camera module; `comp` = the modality chains ({'blocks', 'x', 'i'}) that step along"""
        for blk in blocks:
            live = [c for c in comp if c['i'] < len(c['blocks'])]
            outs = self._run_block_group([blk] + [c['blocks'][c['i']] for c in live], [x] + [c['x'] for c in live])
            x = outs[0]
            for c, o in zip(live, outs[1:]):
                c['x'], c['i'] = o, c['i'] + 1
        return x

    def _can_lockstep_stage(self, mods, comp_stages):
        return (self.ops is ops and hasattr(self.ops, 'window_attention_grouped') and self.precision == 'bf16' and
                os.environ.get('HRF_LOCKSTEP', '1') != '0' and len(comp_stages) > 0 and
                all(rows is None and len(br) == 1 for st in comp_stages for br, rows in st))

    def _run_stage(self, mods, xs, final_nchw=False, comp_stages=None, comp_xs=None):
        """comp_stages / comp_xs: modality stages (single-branch modules) and their inputs to run in
        lockstep with branch 0; returns (xs, [modality outputs]) then"""
        nchw = None
        comp = None
        if comp_stages is not None:
            comp = [dict(blocks=[b for br, _ in st for b in br[0]], x=x, i=0) for st, x in zip(comp_stages, comp_xs)]
        for mi, (branches, rows) in enumerate(mods):
            thunks = [lambda br=br, x=x: self._run_branch(br, x) for br, x in zip(branches, xs)]
            if comp is not None:
                thunks[0] = lambda br=branches[0], x=xs[0]: self._run_branch_lockstep(br, x, comp)
            ys = self._par(thunks)
            if rows is None:
                xs = ys
                continue
            want_nchw = final_nchw and mi == len(mods) - 1
            res = self._par([lambda i=i, u=u, d=d: self._exchange_row(i, u, d, ys, want_nchw)
                             for i, (u, d) in enumerate(rows)])
            if want_nchw:
                xs, nchw = [r[0] for r in res], [r[1] for r in res]
            else:
                xs = res
        if comp is not None:
            for c in comp:                             # blocks beyond the camera's (none in the shipped configs)
                while c['i'] < len(c['blocks']):
                    c['x'], c['i'] = self._run_block(c['blocks'][c['i']], c['x']), c['i'] + 1
            return ((xs, nchw) if final_nchw else xs), [c['x'] for c in comp]
        return (xs, nchw) if final_nchw else xs

    def _conv_group(self, convs, imgs, residuals=None):
        """One layer of the (1 + M) streams.  When every conv runs on the conv-GEMM kernel with
        one layer shape, the streams share ONE launch (hrf_convgemm_grouped_fwd); else each
        stream runs its own."""
        c0 = convs[0]
        same = (len(convs) > 1 and len(convs) <= 4 and all(isinstance(c, _Conv) and c.tc for c in convs) and
                all((c.cout, c.ksize, c.stride1) == (c0.cout, c0.ksize, c0.stride1) for c in convs) and
                all(tuple(i.shape) == tuple(imgs[0].shape) for i in imgs))
        if not same:
            return [c(i) if residuals is None else c(i, r) for c, i, r in
                    zip(convs, imgs, residuals or [None] * len(convs))]
        toks = [self._tokens(i) for i in imgs]
        res = None if residuals is None else [self._tokens(r) for r in residuals]
        outs = ops.conv_gemm_grouped(toks, [c.blob.t for c in convs], c0.cout, c0.ksize, c0.stride1,
                                     [c.relu for c in convs], res)
        return [self._image(o) for o in outs]

    def _stems_lockstep(self, x, mods):
        """The camera stem and the modality stems, layer by layer, with one grouped launch per
        layer; returns (camera tokens per stage-2 branch, modality tokens [k][branch])."""
        chains = [self.stem] + self.stem_mod
        ys = self._par([lambda c=c, t=t: c[0](t) for c, t in zip(chains, [x] + list(mods))])
        ys = self._conv_group([c[1] for c in chains], ys)
        for j in range(2, len(chains[0])):
            bs = [c[j] for c in chains]
            if all(b.cat is not None for b in bs) and all(tuple(y.shape) == tuple(ys[0].shape) for y in ys):
                t = self._conv_group([b.c1 for b in bs], ys)
                t = self._conv_group([b.c2 for b in bs], t)
                ys = run_cat([b.cat for b in bs], t, ys)
                continue
            idt = ys if bs[0].down is None else self._conv_group([b.down for b in bs], ys)
            t = self._conv_group([b.c1 for b in bs], ys)
            t = self._conv_group([b.c2 for b in bs], t)
            ys = self._conv_group([b.c3 for b in bs], t, idt)
        nb_a = len(self.trans1)
        outs = [self._conv_group([self.trans1[i]] + [self.trans_mod['a'][k][i][0] for k in range(self.M)], ys)
                for i in range(nb_a)]
        cams = [self._tokens(outs[i][0]) for i in range(nb_a)]
        pre_ms = [[self._tokens(outs[i][1 + k]) for i in range(nb_a)] for k in range(self.M)]
        return cams, pre_ms

    def _can_lockstep(self):
        if not (self.ops is ops and self.precision == 'bf16' and self.M + 1 <= 4 and
                os.environ.get('HRF_STEM_LOCKSTEP', '1') != '0'):
            return False
        chains = [self.stem] + self.stem_mod
        if any(len(c) != len(chains[0]) for c in chains):
            return False
        for j in range(1, len(chains[0])):
            layer = [c[j] for c in chains]
            if j >= 2 and not all(isinstance(b, _Bottleneck) for b in layer):
                return False
        return all(tr is not None and len(tr) == 1 and isinstance(tr[0], _Conv)
                   for k in range(self.M) for tr in self.trans_mod['a'][k]) and \
            len(self.trans_mod['a'][0]) == len(self.trans1) if self.M else False

    def _apply_chain(self, chain, x_img):
        for c in chain:
            x_img = c(x_img)
        return x_img

    def _fuse(self, letter, cam_thunks, stream, pre_ms=None):
        """cam_thunks[i]() -> camera tokens of branch i.  Returns the fused branches
        and the modality tensors of branch 0 (inputs of the next modality stage).
        `pre_ms[k][i]`: modality tokens already taken through their transition (stage a)."""
        def branch(i):
            ms = []
            for k in range(self.M):
                if pre_ms is not None:
                    ms.append(pre_ms[k][i])
                    continue
                tr = self.trans_mod[letter][k][i]
                ms.append(stream[k] if tr is None else
                          self._tokens(self._apply_chain(tr, self._image(stream[k]))))
            return self._run_block(self.fusion[letter][i], cam_thunks[i](), ms), ms
        res = self._par([lambda i=i: branch(i) for i in range(len(cam_thunks))])
        return [r[0] for r in res], res[0][1]

    def _prep(self, t):
        return t.to(device=self.device, dtype=self.dtype).contiguous(memory_format=torch.channels_last)

    @torch.no_grad()
    def forward(self, x, mods, taps=None):
        """`taps`, if a dict, receives the per-stage feature maps under the names the oracle
        uses (`fusion_a/b/c`, `stage2/3/4`, `stage_b/c`: lists of fp32 NCHW clones) -- north_star's
        "per-stage backbone feature maps" parity is checked on these (tests/test_gpu_backbone.py)."""
        def tap(name, toks):
            if taps is not None:
                taps[name] = [t.permute(0, 3, 1, 2).float().contiguous() for t in toks]
        if len(mods) != self.M:
            raise Exception('num_fused_modalities does not fit the given input length')
        fp32_cuda = self.precision == 'fp32' and self.device.type == 'cuda'
        ctx = torch.backends.cudnn.flags(enabled=True, allow_tf32=False) \
            if fp32_cuda else contextlib.nullcontext()
        dctx = torch.cuda.device(self.device) if self.device.type == 'cuda' \
            else contextlib.nullcontext()
        self._keep = []
        M = self.M
        with ctx, dctx:
            # Stems + first transitions, per stream, in sub-batches of `stem_chunk` frames: the
            # 64 / 256-channel maps of a sub-batch (15.7 MB per 2 frames at 96 x 160) then stay
            # in the 126 MB L2 from the convolution that writes them to the one that reads
            # them; only the narrow 18 / 36-channel transition outputs are concatenated.
            B = x.shape[0]
            ck = self.stem_chunk if 0 < self.stem_chunk < B else B
            nb_a = len(self.trans1)

            def cam_stream():
                parts = []
                for b0 in range(0, B, ck):
                    y = self._apply_chain(self.stem, x[b0:b0 + ck])
                    parts.append([self._tokens(t(y)) for t in self.trans1])
                return parts[0] if len(parts) == 1 else \
                    [torch.cat([q[i] for q in parts]) for i in range(nb_a)]

            def mod_stream(k):
                parts = []
                for b0 in range(0, B, ck):
                    y = self._apply_chain(self.stem_mod[k], mods[k][b0:b0 + ck])
                    outs = []
                    for i in range(nb_a):
                        tr = self.trans_mod['a'][k][i]
                        outs.append(self._tokens(y) if tr is None else
                                    self._tokens(self._apply_chain(tr, y)))
                    parts.append(outs)
                return parts[0] if len(parts) == 1 else \
                    [torch.cat([q[i] for q in parts]) for i in range(nb_a)]

            if self._can_lockstep() and ck == B:
                cams, pre_ms = self._stems_lockstep(x, mods)
            else:
                stems = self._par([cam_stream] + [lambda k=k: mod_stream(k) for k in range(M)])
                cams, pre_ms = stems[0], stems[1:]
            xs, firsts = self._fuse('a', [lambda i=i: cams[i] for i in range(nb_a)], None, pre_ms)
            tap('fusion_a', xs)
            if self._can_lockstep_stage(self.stage[2], self.stage_mod['b']):
                ys, stream = self._run_stage(self.stage[2], xs, comp_stages=self.stage_mod['b'], comp_xs=firsts)
            else:
                res = self._par([lambda: self._run_stage(self.stage[2], xs)] +
                                [lambda k=k: self._run_stage(self.stage_mod['b'][k], [firsts[k]])[0]
                                 for k in range(M)])
                ys, stream = res[0], res[1:]
            tap('stage2', ys)
            tap('stage_b', stream)

            nchw = None
            for idx, letter, nxt in ((2, 'b', 'c'), (3, 'c', 'd')):
                cam_thunks = []
                for i, tr in enumerate(self.trans_cam[idx]):
                    if tr is None:
                        cam_thunks.append(lambda i=i: ys[i])
                    else:
                        cam_thunks.append(lambda tr=tr: self._tokens(
                            self._apply_chain(tr, self._image(ys[-1]))))
                xs, firsts = self._fuse(letter, cam_thunks, stream)
                tap(f'fusion_{letter}', xs)
                last = idx == 3
                cam_stage = self.stage[4 if last else 3]
                if nxt in self.stage_mod and self._can_lockstep_stage(cam_stage, self.stage_mod[nxt]):
                    r0, rest = self._run_stage(cam_stage, xs, final_nchw=last, comp_stages=self.stage_mod[nxt],
                                               comp_xs=firsts)
                    res = [r0] + list(rest)
                else:
                    thunks = [lambda last=last: self._run_stage(cam_stage, xs, final_nchw=last)]
                    if nxt in self.stage_mod:
                        thunks += [lambda k=k: self._run_stage(self.stage_mod[nxt][k], [firsts[k]])[0]
                                   for k in range(M)]
                    res = self._par(thunks)
                if last:
                    ys, nchw = res[0]
                else:
                    ys = res[0]
                tap('stage4' if last else 'stage3', ys)
                if nxt in self.stage_mod:
                    stream = res[1:]
                    tap(f'stage_{nxt}', stream)
            if self.pre_neck_fusion:
                xs, _ = self._fuse('d', [lambda i=i: ys[i] for i in range(len(ys))], stream)
                nchw = [self.ops.fuse_sum(t, relu=True, nchw_out=True)[1] for t in xs]
            self._keep = []
            return nchw


class GraphedForward:
    """One CUDA graph of the whole backbone forward for fixed input tensors.

    The reference is launch-bound (~3 000 kernels per forward, SURVEY.md section
    3.2); replaying a captured graph removes the Python and launch overhead of
    the ~700 launches this engine still issues.  `x` / `mods` are the static
    input buffers: write new data into them (copy_) and call the object."""

    def __init__(self, engine, x, mods, pool=None, warmup=2):
        self.engine, self.x, self.mods = engine, x, list(mods)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                engine.forward(self.x, self.mods)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        n0 = ops.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=pool):
            self.out = engine.forward(self.x, self.mods)
        self.launches = ops.launch_count() - n0      # hrfuser_b200 kernels per replay

    def pool(self):
        return self.graph.pool()

    def __call__(self):
        self.graph.replay()
        return self.out
