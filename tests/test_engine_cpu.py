"""Host logic without a GPU: the C packers + the engine wiring, evaluated through
a blob-level CPU emulation of the device ops (tests/blob_emul.py), must
reproduce the reference goldens; the torch training path must too."""
import copy
import json
import os

import numpy as np
import pytest
import torch

import blob_emul
from helpers import rel_err
from hrfuser_b200 import HRFuserHRFormerBased, backbone_cfg, tiny_cfg
from hrfuser_b200.engine import BackboneEngine
from hrfuser_b200.utils import randomize_parameters, synthetic_inputs

G = os.path.join(os.path.dirname(__file__), 'golden')
E2E = np.load(os.path.join(G, 'e2e.npz'))


def _net(cfg, seed=1):
    c = copy.deepcopy(cfg)
    c.pop('type')
    net = HRFuserHRFormerBased(**c)
    randomize_parameters(net, seed)
    net.eval()
    return net


@pytest.mark.parametrize('tag,v,d,mc', [('t_nus', 't', 'nus', (3, 3)), ('t_stf', 't', 'stf', (3, 2, 1)),
                                        ('b_nus', 'b', 'nus', (3, 3))])
def test_engine_wiring_and_packers_vs_reference_golden(built_lib, tag, v, d, mc):
    net = _net(backbone_cfg(v, d))
    H, W = (int(t) for t in E2E[tag + '_hw'])
    x, mods = synthetic_inputs(1, H, W, mc, seed=3)
    eng = BackboneEngine(net, 'fp32', device_ops=blob_emul)
    out = eng.forward(x, mods)
    assert len(out) == 4
    for i, y in enumerate(out):
        assert y.dtype == torch.float32 and y.is_contiguous()
        assert rel_err(y, torch.from_numpy(E2E[f'{tag}_out{i}'])) < 3e-6


@pytest.mark.parametrize('tag,v,d,mc', [('t_nus', 't', 'nus', (3, 3)), ('t_stf', 't', 'stf', (3, 2, 1))])
def test_training_path_matches_reference_golden(tag, v, d, mc):
    net = _net(backbone_cfg(v, d))
    H, W = (int(t) for t in E2E[tag + '_hw'])
    x, mods = synthetic_inputs(1, H, W, mc, seed=3)
    with torch.no_grad():
        out = net._forward_autograd(x, mods)          # eval-mode BN, torch ops
    for i, y in enumerate(out):
        assert rel_err(y, torch.from_numpy(E2E[f'{tag}_out{i}'])) < 2e-6


def test_state_dict_layout_digest():
    import hashlib
    want = json.load(open(os.path.join(G, 'state_dict_layout.json')))['layouts']
    for tag, (v, d) in {'t_nus': ('t', 'nus'), 't_stf': ('t', 'stf'), 'b_nus': ('b', 'nus')}.items():
        c = backbone_cfg(v, d)
        c.pop('type')
        sd = HRFuserHRFormerBased(**c).state_dict()
        txt = '\n'.join(f'{k} {tuple(t.shape)} {t.dtype}' for k, t in sd.items())
        assert len(sd) == want[tag]['count']
        assert hashlib.sha256(txt.encode()).hexdigest() == want[tag]['digest']


def test_training_mode_backward_and_quirks():
    torch.manual_seed(0)          # DropPath / Dropout masks
    net = _net(tiny_cfg(2))
    net.train()
    x, mods = synthetic_inputs(4, 32, 32, (3, 3), seed=1)
    out = net(x, mods)
    sum(o.sum() for o in out).backward()
    grads = {n: p.grad for n, p in net.named_parameters()}
    # transition1[0] is applied conv-only: its BN never receives a gradient
    assert grads['transition1.0.1.weight'] is None and grads['transition1.0.1.bias'] is None
    assert grads['transition1.0.0.weight'] is not None
    assert grads['fusion_a.0.attn.1.attn.relative_position_bias_table'].abs().sum() > 0
    # drop_path_rate is ignored for HRFormer stages, DropPath(0.2) lives in fusion blocks only
    from hrfuser_b200.modules import DropPath
    dp = [n for n, m in net.named_modules() if isinstance(m, DropPath)]
    assert dp and all(n.startswith('fusion_') for n in dp)


def test_norm_eval_and_extra_mutation():
    cfg = tiny_cfg(2)
    c = copy.deepcopy(cfg)
    c.pop('type')
    c['norm_eval'] = True
    net = HRFuserHRFormerBased(**c)
    net.train()
    assert all(not m.training for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d))
    assert c['extra']['stage3']['drop_path_rates'] == [0.0] * 2
    assert c['extra']['LidarStageC']['drop_path_rates'] is c['extra']['stage3']['drop_path_rates']


def test_bad_configs_raise():
    c = copy.deepcopy(tiny_cfg(2))
    c.pop('type')
    c['extra']['ModFusionA']['block'] = 'NOPE'
    with pytest.raises(Exception):
        HRFuserHRFormerBased(**c)
    c = copy.deepcopy(tiny_cfg(2))
    c.pop('type')
    c['extra']['stage2']['num_channels'] = (18,)
    with pytest.raises(AssertionError):
        HRFuserHRFormerBased(**c)


@pytest.mark.parametrize('tag,v,d,mc', [('t_nus', 't', 'nus', (3, 3)), ('t_stf', 't', 'stf', (3, 2, 1))])
def test_engine_stage_taps_vs_reference_golden(built_lib, tag, v, d, mc):
    """`BackboneEngine.forward(..., taps=...)` names and orders the per-stage maps like the
    reference's forward hooks (tests/golden/e2e.npz), on the blob-level CPU emulation."""
    net = _net(backbone_cfg(v, d))
    H, W = (int(t) for t in E2E[tag + '_hw'])
    x, mods = synthetic_inputs(1, H, W, mc, seed=3)
    eng = BackboneEngine(net, 'fp32', device_ops=blob_emul)
    taps = {}
    eng.forward(x, mods, taps=taps)
    n = 0
    for stage in ('fusion_a', 'fusion_b', 'fusion_c', 'stage2', 'stage3', 'stage4'):
        for i, t in enumerate(taps[stage]):
            assert rel_err(t, torch.from_numpy(E2E[f'{tag}_{stage}.{i}'])) < 3e-6, (stage, i)
            n += 1
    assert n == 2 + 3 + 4 + 2 + 3 + 4
    assert len(taps['stage_b']) == len(mc) and len(taps['stage_c']) == len(mc)


def test_stem_sub_batches_change_nothing(built_lib):
    """HRF_STEM_CHUNK runs the stem phase in sub-batches of frames (an L2-blocking experiment,
    DESIGN.md): frame-wise independent work, so the result is the whole-batch one."""
    net = _net(tiny_cfg(2))
    x, mods = synthetic_inputs(3, 64, 96, (3, 3), seed=8)
    eng = BackboneEngine(net, 'fp32', device_ops=blob_emul)
    a = eng.forward(x, mods)
    for ck in (1, 2):
        eng.stem_chunk = ck
        b = eng.forward(x, mods)
        for p, q in zip(a, b):
            assert rel_err(q, p) < 1e-6


def test_eval_mode_with_trainable_parameters_builds_a_graph():
    """eval() + grad enabled + trainable parameters (fine-tuning with frozen BN
    statistics) must take the autograd path, as the reference does; under no_grad the engine
    runs (and, without a GPU, fails loudly)."""
    net = _net(tiny_cfg(2))
    x, mods = synthetic_inputs(1, 32, 32, (3, 3), seed=2)
    out = net(x, mods)
    assert all(o.requires_grad for o in out)
    sum(o.sum() for o in out).backward()
    assert net.conv1.weight.grad is not None
    from hrfuser_b200 import _lib
    with torch.no_grad(), pytest.raises(_lib.HrfError):
        net(x, mods)


def test_neck_cache_is_dropped_by_train_and_parent_load():
    import torch.nn as nn
    from hrfuser_b200 import HRFPN
    neck = HRFPN([18, 36, 72, 144], 32)
    neck._blobs, neck._blobs_ver = ['stale'], neck._weights_version()
    neck.train()
    assert neck._blobs is None
    neck._blobs = ['stale']
    parent = nn.Module()
    parent.neck = neck
    parent.load_state_dict(parent.state_dict())           # a parent's load reaches _load_from_state_dict
    assert neck._blobs is None
    v0 = neck._weights_version()
    with torch.no_grad():
        neck.reduction_conv.conv.weight.add_(1.0)
    assert neck._weights_version() != v0                   # in-place updates are seen


def test_backbone_engine_is_dropped_by_parent_load():
    import torch.nn as nn
    net = _net(tiny_cfg(2))
    net._engine, net._graphs = object(), {'k': 1}
    parent = nn.Module()
    parent.backbone = net
    parent.load_state_dict(parent.state_dict())
    assert net._engine is None and net._graphs == {}


def test_bottleneck_conv3_downsample_fold_into_one_conv():
    """the algebra behind hrf_convgemm_grouped_cat_fwd: relu(bn3(conv3(y)) + bn_d(downsample(x)))
    (reference resnet.py:287-300) == relu(conv_both(cat[y, x])) with the engine's folded weights"""
    from hrfuser_b200.engine import fold_conv3_downsample
    from hrfuser_b200.modules import make_bottleneck_layer
    torch.manual_seed(3)
    layer = make_bottleneck_layer(64, 64, 2, dict(type='BN'))
    randomize_parameters(layer, 9)
    layer.eval()
    m = layer[0]
    assert m.downsample is not None and layer[1].downsample is None
    both = fold_conv3_downsample(m)
    assert both.in_channels == m.conv3.in_channels + m.downsample[0].in_channels
    x = torch.randn(2, 64, 9, 11)
    with torch.no_grad():
        y = m.relu(m.bn2(m.conv2(m.relu(m.bn1(m.conv1(x))))))
        ref = m.relu(m.bn3(m.conv3(y)) + m.downsample(x))
        got = torch.relu(both(torch.cat([y, x], 1)))
        assert rel_err(got, ref) < 1e-6
        assert rel_err(ref, m(x)) < 1e-6
