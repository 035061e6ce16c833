"""TEST INFRASTRUCTURE ONLY -- CPU restatement (functional torch, fp32) of the reference's
`HRFPN.forward` (mmdet/models/necks/hrfpn.py:77-100) over a state_dict.  Only `tests/` may
import this; the product (`hrfuser_b200/`) never does.

Pinned by tests/golden/neck.npz: outputs of the reference's own `HRFPN` class imported from
the reference tree (tests/golden/make_golden_neck.py; mmcv's `ConvModule` without norm /
activation is a bare `nn.Conv2d` under `.conv`).
"""
import torch
import torch.nn.functional as F


def hrfpn_reduce(sd, inputs):
    """hrfpn.py:79-86: upsample by 2**i, concatenate, 1x1 reduction conv"""
    outs = [inputs[0]]
    for i in range(1, len(inputs)):
        outs.append(F.interpolate(inputs[i], scale_factor=2 ** i, mode='bilinear'))
    out = torch.cat(outs, dim=1)
    return F.conv2d(out, sd['reduction_conv.conv.weight'], sd['reduction_conv.conv.bias'])


def hrfpn_forward(sd, inputs, num_outs=5, pooling='AVG', stride=1):
    """hrfpn.py:77-100"""
    pool = F.max_pool2d if pooling == 'MAX' else F.avg_pool2d
    out = hrfpn_reduce(sd, inputs)
    outs = [out] + [pool(out, kernel_size=2 ** i, stride=2 ** i) for i in range(1, num_outs)]
    return tuple(F.conv2d(outs[i], sd[f'fpn_convs.{i}.conv.weight'], sd[f'fpn_convs.{i}.conv.bias'],
                          stride=stride, padding=1) for i in range(num_outs))
