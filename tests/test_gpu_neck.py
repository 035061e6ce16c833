"""GPU parity of the HRFPN neck (hrfuser_b200/neck.py: four hrf_pw_fwd + one hrf_fuse_sum_fwd for
the upsample / concat / reduction front half) against the reference-made golden and, at the
full HRFuser-T / HRFuser-B sizes, against the oracle."""
import os

import numpy as np
import pytest
import torch

from helpers import assert_parity
from oracle import neck_oracle

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'neck.npz'))
CASES = {'t': ([18, 36, 72, 144], 64), 'b_small': ([78, 156, 312, 624], 16)}


def _net(chans, oc, precision, sd=None, seed=0):
    from hrfuser_b200.neck import HRFPN
    net = HRFPN(in_channels=chans, out_channels=oc, precision=precision).eval()
    if sd is None:
        torch.manual_seed(seed)
        for p in net.parameters():
            torch.nn.init.normal_(p, std=0.05)
    else:
        net.load_state_dict(sd)
    return net


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
@pytest.mark.parametrize('name', sorted(CASES))
def test_against_reference_golden(built_lib, name, mode):
    chans, oc = CASES[name]
    sd = {k[len(name) + 4:]: torch.from_numpy(GOLD[k]) for k in GOLD.files if k.startswith(name + '.sd.')}
    xs = [torch.from_numpy(GOLD[f'{name}.in{i}']).cuda() for i in range(4)]
    net = _net(chans, oc, mode, sd).cuda()
    before = built_lib.hrf_launch_count()
    with torch.no_grad():
        got = net(xs)
    if mode == 'bf16' and oc % 64 == 0:
        # + 4 pooled levels, 5 conv-GEMM launches, 5 exit layouts: the whole neck on the library's kernels
        assert built_lib.hrf_launch_count() - before == 4 + 4 + 1 + 4 + 5 + 5
    else:
        assert built_lib.hrf_launch_count() - before == 4 + 4 + 1 + 1      # layouts, pw, fuse_sum, layout
    assert isinstance(got, tuple) and len(got) == 5
    want = [torch.from_numpy(GOLD[f'{name}.out{i}']) for i in range(5)]
    assert all(g.dtype == torch.float32 for g in got)
    assert_parity(got[0], want[0], mode, f'hrfpn {name} out0')
    # pooled levels average 4..256 zero-mean values: the result is small against its terms, so the
    # error is judged against the scale of the un-pooled map (same bar: 2e-5 fp32, 2e-2 bf16)
    scale = float(want[0].pow(2).mean().sqrt())
    for i in range(1, 5):
        err = float((got[i].cpu() - want[i]).pow(2).mean().sqrt()) / scale
        assert err <= (2e-5 if mode == 'fp32' else 2e-2), (i, err)


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
@pytest.mark.parametrize('chans,grid,B', [
    ([18, 36, 72, 144], (96, 160), 8),       # HRFuser-T nuScenes, 8 frames (cfg2)
    ([18, 36, 72, 144], (96, 312), 2),       # HRFuser-T STF (cfg3)
    ([78, 156, 312, 624], (96, 160), 2),     # HRFuser-B nuScenes (cfg4)
])
def test_reduction_full_size_against_oracle(built_lib, chans, grid, B, mode):
    net = _net(chans, 256, mode, seed=1)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net = net.cuda()
    torch.manual_seed(2)
    xs = [torch.randn(B, c, grid[0] >> i, grid[1] >> i) for i, c in enumerate(chans)]
    want = neck_oracle.hrfpn_reduce(sd, xs)
    with torch.no_grad():
        got = net.reduce([x.cuda() for x in xs])
    assert_parity(got, want, mode, f'hrfpn reduce {chans[0]} {grid}')


def test_whole_neck_full_size_and_zero_input(built_lib):
    chans = [18, 36, 72, 144]
    net = _net(chans, 256, 'fp32', seed=3)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net = net.cuda()
    torch.manual_seed(4)
    xs = [torch.randn(1, c, 96 >> i, 160 >> i) for i, c in enumerate(chans)]
    xs[2].zero_()
    want = neck_oracle.hrfpn_forward(sd, xs)
    with torch.no_grad():
        got = net([x.cuda() for x in xs])
    scale = float(want[0].pow(2).mean().sqrt())
    for i, (g, w) in enumerate(zip(got, want)):
        # the 3x3 convs run in torch on both sides (cuDNN vs CPU); pooled levels: see above
        e = float((g.cpu() - w).pow(2).mean().sqrt()) / scale
        assert e < 2e-5, (i, e)
    with torch.no_grad(), pytest.raises(ValueError):
        net([xs[0].cuda(), xs[1].cuda(), xs[2].cuda(), xs[2].cuda()])


def test_reduction_weights_repacked_after_an_optimizer_step(built_lib):
    """eval -> in-place parameter update -> eval must use the NEW reduction
    weights (the packed blobs are keyed on the parameters' version counters and dropped by
    train())."""
    import torch
    from hrfuser_b200 import HRFPN
    torch.manual_seed(0)
    neck = HRFPN([18, 36, 72, 144], 64, precision='fp32').cuda().eval()
    xs = [torch.randn(1, c, 32 >> i, 48 >> i, device='cuda') for i, c in enumerate([18, 36, 72, 144])]
    with torch.no_grad():
        a = neck(xs)
        ref_a = neck._forward_autograd(xs)
        with torch.no_grad():
            for p_ in neck.parameters():
                p_.add_(0.05 * torch.randn_like(p_))           # what optimizer.step() does
        b = neck(xs)
        ref_b = neck._forward_autograd(xs)
    for g, r in zip(a, ref_a):
        assert float((g - r).norm() / r.norm()) < 2e-3        # (the torch side runs TF32 convs)
    for g, r in zip(b, ref_b):
        assert float((g - r).norm() / r.norm()) < 2e-3, 'stale packed reduction weights'
    assert float((a[0] - b[0]).abs().max()) > 1e-3
    neck.train()
    assert neck._blobs is None


@pytest.mark.parametrize('pooling', ['AVG', 'MAX'])
def test_whole_neck_bf16_on_library_kernels(built_lib, pooling):
    """bf16 mode, 256 output channels: reduction, pooling pyramid (hrf_pool_fwd) and the five
    3x3 convs (hrf_convgemm_fwd) all run on the library's kernels; against the oracle."""
    import torch
    from hrfuser_b200.neck import HRFPN
    chans = [18, 36, 72, 144]
    torch.manual_seed(7)
    net = HRFPN(in_channels=chans, out_channels=256, pooling_type=pooling, precision='bf16').eval()
    for p_ in net.parameters():
        torch.nn.init.normal_(p_, std=0.05)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    xs = [torch.randn(2, c, 96 >> i, 160 >> i) for i, c in enumerate(chans)]
    with torch.no_grad():
        want = net._forward_autograd(xs)                  # the reference's formulation, fp32 CPU
    net = net.cuda()
    before = built_lib.hrf_launch_count()
    with torch.no_grad():
        got = net([x.cuda() for x in xs])
    assert built_lib.hrf_launch_count() - before == 4 + 4 + 1 + 4 + 5 + 5
    scale = float(want[0].pow(2).mean().sqrt())
    for i, (g, w) in enumerate(zip(got, want)):
        assert g.dtype == torch.float32 and g.shape == w.shape
        e = float((g.cpu() - w).pow(2).mean().sqrt()) / scale
        assert e <= 2e-2, (i, e)
    assert_parity(got[0], want[0], 'bf16', f'hrfpn bf16 {pooling} out0')
