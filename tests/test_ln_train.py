"""Train-mode LayerNorm kernels (csrc/ln_train.cuh, hrf_ln_fwd / hrf_ln_bwd) against torch's
LayerNorm in fp64: forward, dx, dgamma, dbeta over every channel width of HRFuser-T / -B, ragged
row counts; the module keeps nn.LayerNorm's interface and CPU behaviour."""
import pytest
import torch
import torch.nn as nn

from hrfuser_b200.bn_train import HrfLayerNorm


def test_module_is_a_layernorm_with_the_same_state_dict_and_cpu_result():
    torch.manual_seed(0)
    m, ref = HrfLayerNorm(36, eps=1e-6), nn.LayerNorm(36, eps=1e-6)
    assert isinstance(m, nn.LayerNorm) and list(m.state_dict()) == list(ref.state_dict())
    with torch.no_grad():
        m.weight.normal_()
        m.bias.normal_()
    ref.load_state_dict(m.state_dict())
    x = torch.randn(3, 50, 36, requires_grad=True)
    y = m(x)
    assert torch.equal(y, ref(x))                       # CPU tensors: torch's implementation
    y.sum().backward()
    assert x.grad is not None


@pytest.mark.gpu
@pytest.mark.parametrize('C', [18, 36, 72, 144, 78, 156, 312, 624, 1000])
@pytest.mark.parametrize('rows', [(2, 96 * 160), (3, 77), (1, 1)])
def test_gpu_forward_backward_match_fp64(built_lib, C, rows):
    from hrfuser_b200 import ops
    B, N = rows
    if C >= 312 and N > 1000:
        N = 1200
    torch.manual_seed(C + N)
    m = HrfLayerNorm(C, eps=1e-6).cuda()
    with torch.no_grad():
        m.weight.copy_(1 + 0.3 * torch.randn(C))
        m.bias.copy_(0.2 * torch.randn(C))
    x = (torch.randn(B, N, C, device='cuda') * 2 + 0.7).requires_grad_()
    dy = torch.randn(B, N, C, device='cuda')
    n0 = ops.launch_count()
    y = m(x)
    y.backward(dy)
    assert ops.launch_count() - n0 == 3                 # forward, backward, finalize
    ref = nn.LayerNorm(C, eps=1e-6).cuda().double()
    ref.load_state_dict({k: v.double() for k, v in m.state_dict().items()})
    xd = x.detach().double().requires_grad_()
    yd = ref(xd)
    yd.backward(dy.double())
    rel = lambda a, b: float((a.double() - b).norm() / b.norm().clamp_min(1e-30))
    assert rel(y, yd) < 2e-6
    assert rel(x.grad, xd.grad) < 5e-6
    assert rel(m.weight.grad, ref.weight.grad) < 2e-5
    assert rel(m.bias.grad, ref.bias.grad) < 2e-5


@pytest.mark.gpu
def test_gpu_backward_is_deterministic_and_skips_dx_when_not_needed(built_lib):
    m = HrfLayerNorm(78, eps=1e-6).cuda()
    x = torch.randn(4, 999, 78, device='cuda')
    dy = torch.randn_like(x)
    grads = []
    for _ in range(2):
        m.zero_grad()
        m(x).backward(dy)                               # x needs no gradient: dx is not computed
        grads.append((m.weight.grad.clone(), m.bias.grad.clone()))
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1])
