"""Generates tests/golden/neck.npz: inputs, state_dict and outputs of the reference's own
`HRFPN` (mmdet/models/necks/hrfpn.py) imported from the read-only reference checkout.

mmcv is not installed: `mmcv.cnn.ConvModule(cin, cout, k, padding, stride, conv_cfg=None,
act_cfg=None)` without a norm layer is an `nn.Conv2d` (bias=True) under the attribute `conv`;
`mmcv.runner.BaseModule` is an `nn.Module` carrying `init_cfg`.  Parameters are seeded-random.

Run from the repo root:  python tests/golden/make_golden_neck.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get('HRFUSER_REF', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_hrfpn():
    class ConvModule(nn.Module):
        def __init__(self, cin, cout, kernel_size, stride=1, padding=0, conv_cfg=None, norm_cfg=None,
                     act_cfg=dict(type='ReLU')):
            super().__init__()
            assert conv_cfg is None and norm_cfg is None and act_cfg is None
            self.conv = nn.Conv2d(cin, cout, kernel_size, stride=stride, padding=padding)

        def forward(self, x):
            return self.conv(x)

    class BaseModule(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()
            self.init_cfg = init_cfg

    class _Reg:
        def register_module(self, *a, **k):
            return lambda cls: cls

    mods = {n: types.ModuleType(n) for n in ('mmcv', 'mmcv.cnn', 'mmcv.runner', 'mmdet', 'mmdet.models',
                                             'mmdet.models.builder', 'mmdet.models.necks')}
    for m in mods.values():
        m.__path__ = []
    mods['mmcv.cnn'].ConvModule = ConvModule
    mods['mmcv.runner'].BaseModule = BaseModule
    mods['mmdet.models.builder'].NECKS = _Reg()
    sys.modules.update(mods)
    spec = importlib.util.spec_from_file_location(
        'mmdet.models.necks.hrfpn', os.path.join(REF, 'mmdet', 'models', 'necks', 'hrfpn.py'))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod.HRFPN


def main():
    HRFPN = load_reference_hrfpn()
    out = {}
    for name, chans, oc, (H, W), B in (('t', [18, 36, 72, 144], 64, (16, 32), 2),
                                       ('b_small', [78, 156, 312, 624], 16, (16, 16), 1)):
        torch.manual_seed(7)
        net = HRFPN(in_channels=chans, out_channels=oc).eval()
        for p in net.parameters():
            nn.init.normal_(p, std=0.05)
        xs = [torch.randn(B, c, H >> i, W >> i) for i, c in enumerate(chans)]
        with torch.no_grad():
            ys = net(xs)
        assert isinstance(ys, tuple) and len(ys) == 5
        for i, x in enumerate(xs):
            out[f'{name}.in{i}'] = x.numpy()
        for i, y in enumerate(ys):
            out[f'{name}.out{i}'] = y.numpy()
        for k, v in net.state_dict().items():
            out[f'{name}.sd.{k}'] = v.numpy()
        print(name, [tuple(y.shape) for y in ys], len(net.state_dict()), 'tensors')
    np.savez_compressed(os.path.join(HERE, 'neck.npz'), **out)


if __name__ == '__main__':
    main()
