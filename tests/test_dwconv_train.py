"""Training-mode depthwise 3x3 conv kernels (csrc/dwconv_train.cuh) against torch's conv2d in fp64:
forward, input gradient, weight / bias gradients, stride 1 and 2, ragged sizes."""
import pytest
import torch
import torch.nn as nn

from hrfuser_b200.bn_train import HrfDepthwiseConv2d, make_conv


def test_factory_and_cpu_behaviour():
    m = make_conv(8, 8, 3, 2, 1, groups=8, bias=False)
    assert isinstance(m, HrfDepthwiseConv2d) and isinstance(m, nn.Conv2d)
    assert type(make_conv(8, 16, 3, 1, 1)) is nn.Conv2d and type(make_conv(8, 8, 1, 1, 0, groups=8)) is nn.Conv2d
    ref = nn.Conv2d(8, 8, 3, 2, 1, groups=8, bias=False)
    ref.load_state_dict(m.state_dict())
    x = torch.randn(2, 8, 9, 11)
    assert torch.equal(m(x), ref(x))                      # CPU tensors: torch's implementation


@pytest.mark.gpu
@pytest.mark.parametrize('stride', [1, 2])
@pytest.mark.parametrize('bias', [False, True])
@pytest.mark.parametrize('shape', [(2, 72, 96, 160), (3, 18, 13, 21), (1, 5, 1, 1), (2, 312, 24, 40), (1, 7, 6, 3)])
def test_gpu_forward_backward_match_fp64(built_lib, stride, bias, shape):
    from hrfuser_b200 import ops
    B, C, H, W = shape
    torch.manual_seed(C + H + stride)
    m = make_conv(C, C, 3, stride, 1, groups=C, bias=bias).cuda()
    ref = nn.Conv2d(C, C, 3, stride, 1, groups=C, bias=bias).cuda().double()
    ref.load_state_dict({k: v.double() for k, v in m.state_dict().items()})
    x = torch.randn(B, C, H, W, device='cuda', requires_grad=True)
    n0 = ops.launch_count()
    y = m(x)
    g = torch.randn_like(y)
    y.backward(g)
    assert ops.launch_count() - n0 == 4                   # forward, dgrad, wgrad + finalize
    xd = x.detach().double().requires_grad_()
    yd = ref(xd)
    yd.backward(g.double())
    rel = lambda a, b: float((a.detach().double() - b).norm() / b.norm().clamp_min(1e-30))
    assert y.shape == yd.shape
    assert rel(y, yd) < 2e-6
    assert rel(x.grad, xd.grad) < 2e-6
    assert rel(m.weight.grad, ref.weight.grad) < 2e-5
    if bias:
        assert rel(m.bias.grad, ref.bias.grad) < 2e-5
