"""Drop-in `HRFuserHRFormerBased` backbone (mmdet operator API).

Boundary mirrored from the reference:
  * registry name / class name `HRFuserHRFormerBased`
    (reference hrfuser_hrformer_based.py:330-331)
  * constructor keywords and `extra` keys (reference :337-351, hrformer.py:652-694,
    hrnet.py:288-407)
  * `forward(x, x_mod)` -> list of 4 NCHW fp32 feature maps (reference :522-627);
    the `forward(img, *modalities)` spelling of BASELINE.json is accepted too
  * `state_dict()` layout -- 2 979 tensors for HRFuser-T nuScenes
  * `train(mode)` honouring `norm_eval` (reference hrnet.py:588-596)

In `eval()` the forward runs on the sm_100a engine (`engine.BackboneEngine`:
hand-written CUDA kernels behind the C-ABI of include/hrfuser_b200.h) and raises
if that library is not built; there is no CPU or eager fallback for inference.
In `train()` the autograd-capable torch forwards of `modules.py` run (batch
statistics / SyncBN / DropPath), as DESIGN.md section "training" explains.
"""
import os

import torch
import torch.nn as nn

from .modules import (HRFormerBlock, HRFormerModule, HRFuserFusionBlock, Bottleneck,
                      make_bottleneck_layer, make_norm, make_transition)
from .bn_train import norm_act


class HRFuserHRFormerBased(nn.Module):
    blocks_dict = {'BOTTLENECK': Bottleneck, 'HRFORMER': HRFormerBlock,
                   'HRFORMERBLOCK': HRFormerBlock,
                   'CA': HRFuserFusionBlock, 'MWCA': HRFuserFusionBlock}

    def __init__(self, extra, in_channels=3, conv_cfg=None,
                 norm_cfg=dict(type='SyncBN', requires_grad=True),
                 transformer_norm_cfg=dict(type='LN', eps=1e-6), norm_eval=False,
                 with_cp=False, drop_path_rate=0., zero_init_residual=False,
                 multiscale_output=True, pretrained=None, init_cfg=None,
                 num_fused_modalities=2, mod_in_channels=[3, 3],
                 precision='fp32'):
        super().__init__()
        if conv_cfg is not None and conv_cfg.get('type', 'Conv2d') not in ('Conv2d', 'Conv'):
            raise NotImplementedError('only plain Conv2d conv_cfg is supported')
        if not (pretrained is None or isinstance(pretrained, str)):
            raise TypeError('pretrained must be a str or None')
        assert not (init_cfg and pretrained), \
            'init_cfg and pretrained cannot be specified at the same time'
        for s in ('stage1', 'stage2', 'stage3', 'stage4'):
            assert s in extra
            assert len(extra[s]['num_blocks']) == extra[s]['num_branches'] == \
                len(extra[s]['num_channels'])
        self.extra = extra
        self.conv_cfg, self.norm_cfg = conv_cfg, norm_cfg
        self.transformer_norm_cfg = transformer_norm_cfg
        self.norm_eval, self.with_cp = norm_eval, with_cp
        self.zero_init_residual = zero_init_residual
        self.pretrained, self.init_cfg = pretrained, init_cfg
        self.num_fused_modalities = M = num_fused_modalities
        self.with_rpe = extra.get('with_rpe', True)
        self.with_pad_mask = extra.get('with_pad_mask', False)
        self.precision = precision
        self._engine = None
        self._graphs = {}

        # The reference computes stochastic-depth rates from a drop_path_rate
        # that the HRFuser subclass never forwards (hrfuser_hrformer_based.py:
        # 352-362), i.e. always from 0: every HRFormer-stage block gets Identity.
        # The injected keys are kept because callers may read them back.
        for s in ('stage2', 'stage3', 'stage4'):
            extra[s]['drop_path_rates'] = [0.0] * (extra[s]['num_blocks'][0] * extra[s]['num_modules'])
        extra['LidarStageB']['drop_path_rates'] = extra['stage2']['drop_path_rates']
        extra['LidarStageC']['drop_path_rates'] = extra['stage3']['drop_path_rates']
        self.pre_neck_fusion = bool(extra.get('LidarStageD'))
        if self.pre_neck_fusion:
            extra['LidarStageD']['drop_path_rates'] = extra['stage4']['drop_path_rates']

        # ---- camera stream (HRNet/HRFormer skeleton) -------------------------
        self.conv1 = nn.Conv2d(in_channels, 64, 3, 2, 1, bias=False)
        self.bn1 = make_norm(norm_cfg, 64)
        self.conv2 = nn.Conv2d(64, 64, 3, 2, 1, bias=False)
        self.bn2 = make_norm(norm_cfg, 64)
        self.relu = nn.ReLU(inplace=True)
        s1 = extra['stage1']
        self.stage1_cfg = s1
        assert s1['block'] == 'BOTTLENECK'
        self.layer1 = make_bottleneck_layer(64, s1['num_channels'][0], s1['num_blocks'][0], norm_cfg)
        pre = [s1['num_channels'][0] * 4]
        for idx in (2, 3, 4):
            cfg = extra[f'stage{idx}']
            setattr(self, f'stage{idx}_cfg', cfg)
            ch = list(cfg['num_channels'])
            setattr(self, f'transition{idx - 1}', make_transition(pre, ch, norm_cfg))
            stage, pre = self._make_stage(
                cfg, ch, multiscale_output if idx == 4 else True)
            setattr(self, f'stage{idx}', stage)

        # ---- extra-modality streams ------------------------------------------
        self.conv_a = nn.ModuleList(nn.Conv2d(mod_in_channels[k], 64, 3, 2, 1, bias=False)
                                    for k in range(M))
        self.norm_a = nn.ModuleList(make_norm(norm_cfg, 64) for _ in range(M))
        self.conv_b = nn.ModuleList(nn.Conv2d(64, 64, 3, 2, 1, bias=False) for _ in range(M))
        self.norm_b = nn.ModuleList(make_norm(norm_cfg, 64) for _ in range(M))
        sa = extra['LidarStageA']
        self.stage_a_cfg = sa
        assert sa['block'] == 'BOTTLENECK'
        self.layer_a = nn.ModuleList(
            make_bottleneck_layer(64, sa['num_channels'][0], sa['num_blocks'][0], norm_cfg)
            for _ in range(M))
        mod_pre = [[sa['num_channels'][0] * 4] for _ in range(M)]
        self._add_fusion('a', mod_pre)
        for letter in ('b', 'c', 'd') if self.pre_neck_fusion else ('b', 'c'):
            # registration order of the reference: stage_x, transition_x, fusion_x
            scfg = extra[f'LidarStage{letter.upper()}']
            setattr(self, f'stage_{letter}_cfg', scfg)
            built = [self._make_stage(scfg, list(scfg['num_channels'])) for _ in range(M)]
            setattr(self, f'stage_{letter}', nn.ModuleList(st for st, _ in built))
            self._add_fusion(letter, [oc for _, oc in built])

    # ------------------------------------------------------------------ builders
    def _make_stage(self, cfg, in_channels, multiscale_output=True):
        if cfg['block'] not in ('HRFORMER', 'HRFORMERBLOCK'):
            raise KeyError(cfg['block'])
        nb, nblocks = cfg['num_branches'], cfg['num_blocks']
        dpr = cfg['drop_path_rates']
        mods = []
        for i in range(cfg['num_modules']):
            mods.append(HRFormerModule(
                nb, nblocks, in_channels, cfg['num_heads'], cfg['window_sizes'],
                cfg['mlp_ratios'],
                multiscale_output or i != cfg['num_modules'] - 1,
                self.norm_cfg, self.transformer_norm_cfg,
                drop_paths=dpr[nblocks[0] * i:nblocks[0] * (i + 1)],
                with_rpe=self.with_rpe, with_pad_mask=self.with_pad_mask))
        return nn.Sequential(*mods), list(in_channels)

    def _add_fusion(self, letter, mod_pre):
        fcfg = self.extra[f'ModFusion{letter.upper()}']
        setattr(self, f'fusion_{letter}_cfg', fcfg)
        ch = list(fcfg['num_channels'])
        setattr(self, f'transition_{letter}', nn.ModuleList(
            make_transition(mod_pre[k], ch, self.norm_cfg)
            for k in range(self.num_fused_modalities)))
        setattr(self, f'fusion_{letter}', self._make_fusion(fcfg, ch))

    def _make_fusion(self, cfg, channels):
        if cfg['block'] not in ('CA', 'MWCA'):
            raise Exception('Not valid fusion block')
        return nn.ModuleList(
            HRFuserFusionBlock(channels[b], cfg['num_channels'][b],
                               num_heads=cfg['num_heads'][b],
                               window_size=cfg['window_sizes'][b],
                               mlp_ratio=cfg['mlp_ratios'][b], drop_path=cfg['drop_path'],
                               norm_cfg=self.norm_cfg,
                               transformer_norm_cfg=self.transformer_norm_cfg,
                               num_fused_modalities=self.num_fused_modalities,
                               proj_drop_rate=cfg['proj_drop_rate'])
            for b in range(cfg['num_branches']))

    # ------------------------------------------------------------------ protocol
    @property
    def norm1(self):
        return self.bn1

    @property
    def norm2(self):
        return self.bn2

    def init_weights(self):
        """Kaiming for convs, 1/0 for norm layers, trunc-normal(0.02) for the
        relative-position tables (what the reference's init_cfg + per-module
        init_weights amount to); `pretrained`/`init_cfg=Pretrained` checkpoints
        are loaded with torch.load."""
        ckpt = self.pretrained
        if isinstance(self.init_cfg, dict) and self.init_cfg.get('type') == 'Pretrained':
            ckpt = self.init_cfg['checkpoint']
        if ckpt:
            sd = torch.load(ckpt, map_location='cpu')
            sd = sd.get('state_dict', sd)
            sd = {k[len('backbone.'):] if k.startswith('backbone.') else k: v for k, v in sd.items()}
            self.load_state_dict(sd, strict=False)
            return
        for name, m in self.named_modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, (nn.modules.batchnorm._BatchNorm, nn.GroupNorm)):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        for name, p in self.named_parameters():
            if name.endswith('relative_position_bias_table'):
                nn.init.trunc_normal_(p, std=0.02)
        self.invalidate_engine()

    def train(self, mode=True):
        super().train(mode)
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, nn.modules.batchnorm._BatchNorm):
                    m.eval()
        if mode:
            self.invalidate_engine()
        return self

    def invalidate_engine(self):
        """Drop the packed device weights and the captured CUDA graphs (call after mutating
        parameters in place)."""
        self._engine = None
        self._graphs = {}

    def _apply(self, fn, *a, **k):
        self.invalidate_engine()
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *a, **k):
        # reached for every module of the tree when a PARENT (the detector, mmcv's
        # load_checkpoint) loads a checkpoint; Module.load_state_dict of a child is not
        self.invalidate_engine()
        return super()._load_from_state_dict(*a, **k)

    def engine(self):
        if self._engine is None:
            from .engine import BackboneEngine
            self._engine = BackboneEngine(self, precision=self.precision)
        return self._engine

    # ------------------------------------------------------------------ forward
    def forward(self, x, x_mod=None, *more):
        if isinstance(x_mod, torch.Tensor):          # forward(img, lidar, radar, ...)
            x_mod = [x_mod, *more]
        elif x_mod is None:
            x_mod = []
        if self.num_fused_modalities != len(x_mod):
            raise Exception('num_fused_modalities does not fit the given input length')
        if self.training or torch.is_grad_enabled() and (
                any(t.requires_grad for t in (x, *x_mod)) or
                any(p.requires_grad for p in self.parameters())):
            # (eval() with trainable parameters and grad enabled -- fine-tuning with frozen BN
            # statistics -- builds an autograd graph, as the reference does; inference runs
            # under torch.no_grad(), as mmdet's test loop does)
            return self._forward_autograd(x, list(x_mod))
        return self._forward_engine(x, list(x_mod))

    # CUDA-graph cache of the inference forward: the second call with a given input signature
    # captures the whole engine forward into one graph over static input buffers; later calls
    # copy the inputs in (host tensors go straight from pinned memory), replay, and return
    # clones of the four maps (`graph_outputs='view'` returns the static buffers themselves:
    # valid until the next forward).  HRF_GRAPH=0 or `use_cuda_graph = False` keeps the
    # eager launch sequence.
    use_cuda_graph = os.environ.get('HRF_GRAPH', '1') != '0'
    graph_outputs = 'clone'
    max_graphs = 8

    def _forward_engine(self, x, mods):
        eng = self.engine()
        if not (self.use_cuda_graph and eng.device.type == 'cuda') or torch.cuda.is_current_stream_capturing():
            return eng.forward(x, mods)
        # one graph (static buffers, memory pool) per input signature AND stream: a caller that
        # pipelines steps over several streams gets independent instances
        key = (tuple(x.shape), tuple(tuple(m.shape) for m in mods),
               torch.cuda.current_stream(eng.device).cuda_stream)
        ent = self._graphs.get(key)
        if ent is None:                              # first sight: run eagerly, remember
            if len(self._graphs) >= self.max_graphs:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = 0
            return eng.forward(x, mods)
        if ent == 0:
            from .engine import GraphedForward
            with torch.cuda.device(eng.device):
                sx = torch.empty(x.shape, dtype=torch.float32, device=eng.device)
                sm = [torch.empty(m.shape, dtype=torch.float32, device=eng.device) for m in mods]
                sx.copy_(x, non_blocking=True)
                for d, m in zip(sm, mods):
                    d.copy_(m, non_blocking=True)
                ent = self._graphs[key] = GraphedForward(eng, sx, sm)
        else:
            ent.x.copy_(x, non_blocking=True)
            for d, m in zip(ent.mods, mods):
                d.copy_(m, non_blocking=True)
        outs = ent()
        return list(outs) if self.graph_outputs == 'view' else [o.clone() for o in outs]

    def _forward_autograd(self, x, mods):
        """Training path (torch ops, autograd).  Same wiring as the engine; see
        the reference forward hrfuser_hrformer_based.py:522-627."""
        M = self.num_fused_modalities
        x = norm_act(self.bn1, self.relu, self.conv1(x))
        x = self.layer1(norm_act(self.bn2, self.relu, self.conv2(x)))
        stream = []
        for k in range(M):
            m = norm_act(self.norm_a[k], self.relu, self.conv_a[k](mods[k]))
            stream.append(self.layer_a[k](norm_act(self.norm_b[k], self.relu, self.conv_b[k](m))))

        def fuse(letter, cams):
            trans, fusion = getattr(self, f'transition_{letter}'), getattr(self, f'fusion_{letter}')
            outs, firsts = [], None
            for i, cam in enumerate(cams):
                ms = [stream[k] if trans[k][i] is None else trans[k][i](stream[k]) for k in range(M)]
                if i == 0:
                    firsts = ms
                outs.append(fusion[i](cam, ms))
            return outs, firsts

        # stage 2: note transition1[i][0] -- on branch 0 that is the bare conv
        # (no BN / ReLU), on branch 1 the whole conv-bn-relu (reference :550-551)
        cams = [self.transition1[i][0](x) for i in range(self.stage2_cfg['num_branches'])]
        xs, firsts = fuse('a', cams)
        ys = self.stage2(xs)
        stream = [self.stage_b[k]([firsts[k]])[0] for k in range(M)]

        cams = [ys[i] if t is None else t(ys[-1]) for i, t in enumerate(self.transition2)]
        xs, firsts = fuse('b', cams)
        ys = self.stage3(xs)
        stream = [self.stage_c[k]([firsts[k]])[0] for k in range(M)]

        cams = [ys[i] if t is None else t(ys[-1]) for i, t in enumerate(self.transition3)]
        xs, firsts = fuse('c', cams)
        ys = self.stage4(xs)
        if self.pre_neck_fusion:
            stream = [self.stage_d[k]([firsts[k]])[0] for k in range(M)]
            xs, _ = fuse('d', ys)
            ys = [self.relu(t) for t in xs]
        return ys


def register_with_mmdet():
    """Register the class under its reference name when mmdet is importable."""
    try:
        from mmdet.models.builder import BACKBONES
    except Exception:
        return False
    BACKBONES.register_module(name='HRFuserHRFormerBased', force=True,
                              module=HRFuserHRFormerBased)
    return True
