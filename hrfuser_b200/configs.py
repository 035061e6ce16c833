"""Backbone config dictionaries for the shipped HRFuser variants.

These reproduce the *values* of `model.backbone` that the reference's config
tree resolves to (reference: configs/_base_/models/cascade_rcnn_hrfuser_fpn_
{nus_clr,stf_clrg}_fusion.py:7-127 merged with configs/hrfuser/*.py), so that
the drop-in backbone can be built without mmcv's Config machinery.  A dict read
from the reference's own config files by mmcv works the same way: the backbone
takes exactly these keys.
"""
import copy

_WIDTHS = {
    't': dict(channels=(18, 36, 72, 144), heads=(1, 2, 4, 8), stage3_modules=3,
              stagec_modules=3, drop_path_rate=0.0),
    'b': dict(channels=(78, 156, 312, 624), heads=(2, 4, 8, 16), stage3_modules=4,
              stagec_modules=4, drop_path_rate=0.4),
}


def _hrformer_stage(nb, ch, heads, modules, fusion_flag=True):
    d = dict(num_modules=modules, num_branches=nb, block='HRFORMER',
             window_sizes=(7,) * nb, num_heads=tuple(heads[:nb]),
             mlp_ratios=(4,) * nb, num_blocks=(2,) * nb,
             num_channels=tuple(ch[:nb]))
    if fusion_flag:
        d['in_module_fusion'] = False      # present in the reference configs, ignored
    return d


def _fusion(nb, ch, heads, proj_drop_rate=0.1):
    return dict(block='MWCA', with_act=True, with_pre_act=False, drop_path=0.2,
                num_branches=nb, window_sizes=(7,) * nb,
                num_heads=tuple(heads[:nb]), mlp_ratios=(4,) * nb,
                num_channels=tuple(ch[:nb]), proj_drop_rate=proj_drop_rate)


def backbone_cfg(variant='t', dataset='nus', norm='BN'):
    """variant: 't' | 'b';  dataset: 'nus' (cam+lidar+radar) | 'stf'
    (cam+lidar+radar+gated);  norm: 'BN' (the *_bn configs) | 'SyncBN'."""
    w = _WIDTHS[variant]
    ch, heads = w['channels'], w['heads']
    bottleneck = dict(num_modules=1, num_branches=1, block='BOTTLENECK',
                      num_blocks=(2,), num_channels=(64,))
    extra = dict(
        LidarStageA=copy.deepcopy(bottleneck),
        ModFusionA=_fusion(2, ch, heads),
        LidarStageB=_hrformer_stage(1, ch, heads, 1, fusion_flag=False),
        ModFusionB=_fusion(3, ch, heads),
        LidarStageC=_hrformer_stage(1, ch, heads, w['stagec_modules'], fusion_flag=False),
        ModFusionC=_fusion(4, ch, heads),
        LidarStageD=None,
        stage1=copy.deepcopy(bottleneck),
        stage2=_hrformer_stage(2, ch, heads, 1),
        stage3=_hrformer_stage(3, ch, heads, w['stage3_modules']),
        stage4=_hrformer_stage(4, ch, heads, 2))
    cfg = dict(type='HRFuserHRFormerBased',
               norm_cfg=dict(type=norm, requires_grad=True, momentum=0.1),
               transformer_norm_cfg=dict(type='LN', eps=1e-6),
               norm_eval=False, drop_path_rate=w['drop_path_rate'],
               num_fused_modalities=2, extra=extra)
    if dataset == 'stf':
        cfg['num_fused_modalities'] = 3
        cfg['mod_in_channels'] = [3, 2, 1]
    elif dataset != 'nus':
        raise KeyError(dataset)
    return cfg


# name -> (variant, dataset, H, W): the padded network input sizes of the
# reference's pipelines
WORKLOADS = {
    'hrfuser_t_nus_r640': ('t', 'nus', 384, 640),
    'hrfuser_t_stf_r1248': ('t', 'stf', 384, 1248),
    # BASELINE.json's literal "1248x666" (un-cropped STF frame padded to /32); the shipped
    # pipeline crops to 384x1248 (kitti_detection_2d_c1248_clrg_fusion.py:24,55)
    'hrfuser_t_stf_r1248_full': ('t', 'stf', 672, 1248),
    'hrfuser_b_nus_r640': ('b', 'nus', 384, 640),
}


def tiny_cfg(num_mod=2, mod_in_channels=None, channels=(18, 36, 72, 144),
             heads=(1, 2, 4, 8), modules=(1, 1, 1)):
    """A shrunken topology (one module per stage) for fast tests."""
    cfg = backbone_cfg('t', 'nus')
    e = cfg['extra']
    for key, nb in (('ModFusionA', 2), ('ModFusionB', 3), ('ModFusionC', 4)):
        e[key]['num_channels'] = tuple(channels[:nb])
        e[key]['num_heads'] = tuple(heads[:nb])
    for key in ('LidarStageB', 'LidarStageC'):
        e[key]['num_channels'] = (channels[0],)
        e[key]['num_heads'] = (heads[0],)
    e['LidarStageC']['num_modules'] = modules[1]
    for i, (key, nb) in enumerate((('stage2', 2), ('stage3', 3), ('stage4', 4))):
        e[key]['num_channels'] = tuple(channels[:nb])
        e[key]['num_heads'] = tuple(heads[:nb])
        e[key]['num_modules'] = modules[i]
    cfg['num_fused_modalities'] = num_mod
    cfg['mod_in_channels'] = list(mod_in_channels or [3] * num_mod)
    return cfg
