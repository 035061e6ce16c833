"""HRFPN neck, CPU side: oracle and training path against the reference-made
golden, state_dict layout, and the algebra the GPU path rests on (1x1 conv commutes with
bilinear upsampling), checked in fp64 torch."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import neck_oracle
from hrfuser_b200 import _lib
from hrfuser_b200.neck import HRFPN

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'neck.npz'))
CASES = {'t': ([18, 36, 72, 144], 64), 'b_small': ([78, 156, 312, 624], 16)}


def case(name):
    sd = {k[len(name) + 4:]: torch.from_numpy(GOLD[k]) for k in GOLD.files if k.startswith(name + '.sd.')}
    xs = [torch.from_numpy(GOLD[f'{name}.in{i}']) for i in range(4)]
    ys = [torch.from_numpy(GOLD[f'{name}.out{i}']) for i in range(5)]
    return sd, xs, ys


@pytest.mark.parametrize('name', sorted(CASES))
def test_oracle_against_reference_golden(name):
    sd, xs, want = case(name)
    got = neck_oracle.hrfpn_forward(sd, xs)
    assert len(got) == 5
    for g, w in zip(got, want):
        assert g.shape == w.shape
        assert float((g - w).norm() / w.norm()) < 1e-6


@pytest.mark.parametrize('name', sorted(CASES))
def test_state_dict_layout_and_training_path(name):
    chans, oc = CASES[name]
    sd, xs, want = case(name)
    net = HRFPN(in_channels=chans, out_channels=oc)
    assert list(net.state_dict().keys()) == list(sd.keys())          # reference key order
    net.load_state_dict(sd)
    got = net.train()(xs)                                            # torch-autograd path
    assert isinstance(got, tuple)
    for g, w in zip(got, want):
        assert float((g - w).norm() / w.norm()) < 1e-6
    got[0].sum().backward()
    assert net.reduction_conv.conv.weight.grad is not None


def test_reduction_commutes_with_upsampling_fp64():
    sd, xs, _ = case('t')
    sd = {k: v.double() for k, v in sd.items()}
    xs = [x.double() for x in xs]
    want = neck_oracle.hrfpn_reduce(sd, xs)
    w, b = sd['reduction_conv.conv.weight'], sd['reduction_conv.conv.bias']
    off, acc = 0, None
    for i, x in enumerate(xs):
        y = F.conv2d(x, w[:, off:off + x.shape[1]], b if i == 0 else None)
        if i:
            y = F.interpolate(y, scale_factor=2 ** i, mode='bilinear')
        acc = y if acc is None else acc + y
        off += x.shape[1]
    assert float((acc - want).abs().max()) < 1e-12


def test_eval_forward_without_gpu_fails_loudly():
    chans, oc = CASES['t']
    sd, xs, _ = case('t')
    net = HRFPN(in_channels=chans, out_channels=oc).eval()
    net.load_state_dict(sd)
    with torch.no_grad(), pytest.raises(_lib.HrfError):
        net(xs)


def test_ctor_contract():
    with pytest.raises(AssertionError):
        HRFPN(in_channels=18, out_channels=8)                        # hrfpn.py:49
    with pytest.raises(AssertionError):
        HRFPN(in_channels=[18, 36], out_channels=8).train()([torch.zeros(1, 18, 4, 4)])   # hrfpn.py:78


def test_backbone_feeds_neck_training_path():
    """extract_feat of the reference (two_stage.py:76-84): neck(backbone(img, mods)) on the
    torch-autograd path, gradients reaching the backbone through the neck"""
    import copy
    from hrfuser_b200 import HRFuserHRFormerBased, tiny_cfg
    c = copy.deepcopy(tiny_cfg(2))
    c.pop('type')
    c['norm_cfg'] = dict(type='BN', requires_grad=True)
    torch.manual_seed(0)
    backbone = HRFuserHRFormerBased(**c).train()
    neck = HRFPN(in_channels=[18, 36, 72, 144], out_channels=16).train()
    neck.init_weights()
    img, mods = torch.randn(1, 3, 64, 64), [torch.randn(1, 3, 64, 64) for _ in range(2)]
    feats = backbone(img, mods)
    assert [f.shape[1] for f in feats] == [18, 36, 72, 144]
    outs = neck(feats)
    assert [tuple(o.shape) for o in outs] == [(1, 16, 16 >> i, 16 >> i) for i in range(5)]
    sum(o.sum() for o in outs).backward()
    assert backbone.conv1.weight.grad is not None and neck.fpn_convs[4].conv.weight.grad is not None
