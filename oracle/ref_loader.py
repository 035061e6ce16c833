"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference backbone.

Imports the four hot-path files of the reference
(`mmdet/models/backbones/{hrfuser_hrformer_based,hrformer,hrnet,resnet}.py`)
from a read-only checkout under a minimal stand-in for `mmcv` (mmcv-full 1.3.17
is pinned by the reference's README.md:41 but is not installed here and there
is no network).  On this path mmcv contributes *constructors only*
(`build_conv_layer` -> nn.Conv2d, `build_norm_layer` -> BatchNorm2d /
SyncBatchNorm / LayerNorm, `build_activation_layer` -> nn.GELU / nn.ReLU,
`build_dropout` -> DropPath, BaseModule/ModuleList/Sequential); all arithmetic
is torch ATen.  Nothing from the reference is copied: its files are imported
from where they lie.

Used by
  * `tests/golden/make_golden.py`   (generates the committed golden vectors)
  * `tests/test_oracle_vs_reference.py` (pins `oracle/hrfuser_oracle.py`;
     skipped when no reference checkout is reachable, e.g. on the GPU box)

It must never be imported by the product package `hrfuser_b200/`.
"""
import ast
import importlib
import os
import sys
import types
import warnings

import torch
import torch.nn as nn

_SEARCH = [os.environ.get('HRFUSER_REF', ''), '/root/reference',
           os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        'baseline', '_ref')]


def find_reference():
    for p in _SEARCH:
        if p and os.path.isfile(os.path.join(
                p, 'mmdet', 'models', 'backbones', 'hrfuser_hrformer_based.py')):
            return p
    return None


# ----------------------------------------------------------------------------
# mmcv stand-in
# ----------------------------------------------------------------------------
class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg

    def init_weights(self):
        for m in self.children():
            if hasattr(m, 'init_weights'):
                m.init_weights()


class _ModuleList(_BaseModule, nn.ModuleList):
    def __init__(self, modules=None, init_cfg=None):
        _BaseModule.__init__(self, init_cfg)
        nn.ModuleList.__init__(self, modules)


class _Sequential(_BaseModule, nn.Sequential):
    def __init__(self, *args, init_cfg=None):
        _BaseModule.__init__(self, init_cfg)
        nn.Sequential.__init__(self, *args)


class _DropPath(nn.Module):
    """Stochastic depth per sample (mmcv.cnn.bricks.drop.DropPath)."""

    def __init__(self, drop_prob=0.1):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0. or not self.training:
            return x
        keep = 1 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = keep + torch.rand(shape, dtype=x.dtype, device=x.device)
        return x.div(keep) * mask.floor()


def _build_conv_layer(cfg, *args, **kwargs):
    assert cfg is None or cfg.get('type', 'Conv2d') in ('Conv2d', 'Conv')
    return nn.Conv2d(*args, **kwargs)


def _build_norm_layer(cfg, num_features, postfix=''):
    cfg = dict(cfg)
    t = cfg.pop('type')
    requires_grad = cfg.pop('requires_grad', True)
    cfg.setdefault('eps', 1e-5)
    if t in ('BN', 'BN2d'):
        layer, abbr = nn.BatchNorm2d(num_features, **cfg), 'bn'
    elif t == 'SyncBN':
        layer, abbr = nn.SyncBatchNorm(num_features, **cfg), 'bn'
    elif t == 'LN':
        layer, abbr = nn.LayerNorm(num_features, **cfg), 'ln'
    else:
        raise KeyError(t)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return abbr + str(postfix), layer


def _build_activation_layer(cfg):
    t = cfg['type']
    if t == 'GELU':
        return nn.GELU()
    if t == 'ReLU':
        return nn.ReLU(inplace=cfg.get('inplace', False))
    raise KeyError(t)


def _build_dropout(cfg):
    cfg = dict(cfg)
    t = cfg.pop('type')
    if t == 'DropPath':
        return _DropPath(**cfg)
    if t == 'Dropout':
        return nn.Dropout(cfg.get('drop_prob', cfg.get('p', 0.5)))
    raise KeyError(t)


def _constant_init(module, val, bias=0):
    if hasattr(module, 'weight') and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def _trunc_normal_init(module, mean=0., std=1., a=-2., b=2., bias=0.):
    if hasattr(module, 'weight') and module.weight is not None:
        nn.init.trunc_normal_(module.weight, mean, std, a, b)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


class _Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def get(self, key):
        return self.module_dict.get(key)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_LOADED = {}


def load_reference(ref_root=None):
    """Return the reference module `mmdet.models.backbones.hrfuser_hrformer_based`."""
    ref_root = ref_root or find_reference()
    if ref_root is None:
        raise FileNotFoundError('no reference checkout (set $HRFUSER_REF)')
    if ref_root in _LOADED:
        return _LOADED[ref_root]
    if 'mmcv' in sys.modules and not getattr(sys.modules['mmcv'], '_hrf_standin', False):
        raise RuntimeError('a real mmcv is already imported; the stand-in would clash')

    # -- mmcv ------------------------------------------------------------
    mmcv = _mod('mmcv', __version__='1.3.17', _hrf_standin=True)
    mmcv.__path__ = []
    cnn = _mod('mmcv.cnn', build_conv_layer=_build_conv_layer,
               build_norm_layer=_build_norm_layer,
               build_activation_layer=_build_activation_layer,
               build_plugin_layer=lambda *a, **k: (_ for _ in ()).throw(NotImplementedError()),
               constant_init=_constant_init, trunc_normal_init=_trunc_normal_init,
               MODELS=_Registry('model'))
    cnn.__path__ = []
    bricks = _mod('mmcv.cnn.bricks')
    bricks.__path__ = []
    _mod('mmcv.cnn.bricks.transformer', build_dropout=_build_dropout)
    _mod('mmcv.runner', BaseModule=_BaseModule, ModuleList=_ModuleList,
         Sequential=_Sequential, _load_checkpoint=lambda *a, **k: {})
    utils = _mod('mmcv.utils')
    utils.__path__ = []
    _mod('mmcv.utils.parrots_wrapper', _BatchNorm=nn.modules.batchnorm._BatchNorm)

    # -- bare mmdet packages whose __path__ points into the reference ------
    def pkg(name, *rel):
        m = _mod(name)
        m.__path__ = [os.path.join(ref_root, *rel)]
        return m
    pkg('mmdet', 'mmdet')
    _mod('mmdet.utils', get_root_logger=lambda *a, **k: __import__('logging').getLogger('mmdet'))
    pkg('mmdet.models', 'mmdet', 'models')
    _mod('mmdet.models.builder', BACKBONES=_Registry('backbone'))
    pkg('mmdet.models.backbones', 'mmdet', 'models', 'backbones')

    # layout helpers: run the reference's own FunctionDefs, extracted with ast
    src_path = os.path.join(ref_root, 'mmdet', 'models', 'utils', 'transformer.py')
    with open(src_path) as f:
        tree = ast.parse(f.read())
    want = {'nlc_to_nchw', 'nchw_to_nlc', 'nlc2nchw2nlc'}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    assert {n.name for n in body} == want
    ns = {'torch': torch, 'nn': nn}
    exec(compile(ast.Module(body=body, type_ignores=[]), src_path, 'exec'), ns)

    class _ResLayerStub(nn.Sequential):
        pass
    _mod('mmdet.models.utils', ResLayer=_ResLayerStub,
         **{k: ns[k] for k in want})

    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ref = importlib.import_module('mmdet.models.backbones.hrfuser_hrformer_based')
    _LOADED[ref_root] = ref
    return ref


def build_reference_backbone(backbone_cfg, ref_root=None):
    """Instantiate the reference `HRFuserHRFormerBased` from a backbone dict
    (the value of `model.backbone` in the reference's configs)."""
    import copy
    ref = load_reference(ref_root)
    cfg = copy.deepcopy(backbone_cfg)
    cfg.pop('type', None)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        net = ref.HRFuserHRFormerBased(**cfg)
    net.eval()          # NB: returns None in the reference (hrnet.py:588-596)
    return net


# ----------------------------------------------------------------------------
# reference config reader (`_base_` deep merge honouring `_delete_`)
# ----------------------------------------------------------------------------
def _merge(base, over):
    out = dict(base)
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get('_delete_', False):
            out[k] = _merge(out[k], v)
        else:
            if isinstance(v, dict):
                v = {kk: vv for kk, vv in v.items() if kk != '_delete_'}
            out[k] = v
    return out


def read_reference_config(rel_path, ref_root=None):
    ref_root = ref_root or find_reference()
    path = os.path.join(ref_root, rel_path)

    def load(p):
        ns = {}
        with open(p) as f:
            exec(compile(f.read(), p, 'exec'), ns)
        cfg = {k: v for k, v in ns.items()
               if not k.startswith('__') and not isinstance(v, types.ModuleType)}
        bases = cfg.pop('_base_', [])
        if isinstance(bases, str):
            bases = [bases]
        merged = {}
        for b in bases:
            merged = _merge(merged, load(os.path.normpath(os.path.join(os.path.dirname(p), b))))
        return _merge(merged, cfg)
    return load(path)
