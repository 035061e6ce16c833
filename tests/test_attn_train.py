"""Training-mode window-attention core (csrc/attn_train.cuh) against the torch formulation of the
module (`_WindowAttnBase._core`) in fp64: forward, dq / dk / dv and the relative-position table
gradient, for the head dims / head counts of HRFuser-T and -B, self and cross."""
import pytest
import torch

from hrfuser_b200.modules import WindowMCA, WindowMSA

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('C,heads', [(18, 1), (36, 2), (144, 8), (78, 2), (156, 4), (624, 16)])
@pytest.mark.parametrize('cross', [False, True])
def test_core_forward_backward_match_fp64(built_lib, C, heads, cross):
    from hrfuser_b200 import ops
    torch.manual_seed(C + heads)
    nWin = 37 if C < 300 else 5
    cls = WindowMCA if cross else WindowMSA
    m = cls(C, heads, (7, 7)).cuda().train()
    with torch.no_grad():
        m.relative_position_bias_table.normal_(0, 0.5)
    ref = cls(C, heads, (7, 7)).cuda().double().train()
    ref.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in m.state_dict().items()})
    x = torch.randn(nWin, 49, C, device='cuda', requires_grad=True)
    z = torch.randn(nWin, 49, C, device='cuda', requires_grad=True)
    dy = torch.randn(nWin, 49, C, device='cuda')
    n0 = ops.launch_count()
    y = m(x, z if cross else None)
    y.backward(dy)
    assert ops.launch_count() - n0 == 4               # forward, backward, table gradient (sum + fold)
    xd, zd = x.detach().double().requires_grad_(), z.detach().double().requires_grad_()
    ref._core_force_torch = True
    yd = ref(xd, zd if cross else None)               # fp64 tensors take the torch formulation
    yd.backward(dy.double())
    rel = lambda a, b: float((a.double() - b).norm() / b.norm().clamp_min(1e-30))
    assert rel(y, yd) < 5e-6
    assert rel(x.grad, xd.grad) < 2e-5
    if cross:
        assert rel(z.grad, zd.grad) < 2e-5
    assert rel(m.relative_position_bias_table.grad, ref.relative_position_bias_table.grad) < 2e-5
    for (n, p), (_, pr) in zip(m.named_parameters(), ref.named_parameters()):
        if n == 'k_proj.bias':      # a key bias shifts every logit of a row alike: its gradient is exactly 0
            assert float(p.grad.abs().max()) < 1e-3 * float(dy.abs().max()), n
        else:
            assert rel(p.grad, pr.grad) < 5e-5, n


def test_table_gradient_is_deterministic(built_lib):
    torch.manual_seed(3)
    m = WindowMSA(78, 2, (7, 7)).cuda().train()
    x = torch.randn(644, 49, 78, device='cuda')
    dy = torch.randn_like(x)
    gs = []
    for _ in range(2):
        m.zero_grad()
        m(x).backward(dy)
        gs.append(m.relative_position_bias_table.grad.clone())
    assert torch.equal(gs[0], gs[1])
