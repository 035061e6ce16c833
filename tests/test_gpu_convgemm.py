"""GPU parity of the warp-specialised TMA / tcgen05 conv-GEMM kernel (hrf_convgemm_fwd: stem
conv2, Bottleneck convs, 256-channel transitions, HRFPN 3x3 convs -- SURVEY.md 8f rank 1 / 2)
against torch's fp32 convolution of the same bf16-rounded operands."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from helpers import assert_parity
from hrfuser_b200.utils import randomize_parameters

pytestmark = pytest.mark.gpu

# (Cin, Cout, k, stride, relu, residual)
LAYERS = [
    (64, 64, 1, 1, True, False),       # Bottleneck conv1 of the first block (64 -> 64)
    (256, 64, 1, 1, True, False),      # Bottleneck conv1 (256 -> 64)
    (64, 64, 3, 1, True, False),       # Bottleneck conv2
    (64, 256, 1, 1, True, True),       # Bottleneck conv3 + identity + ReLU
    (64, 256, 1, 1, False, False),     # downsample (bias-free in the engine, here with BN)
    (64, 64, 3, 2, True, False),       # stem conv2 (stride 2)
    (256, 18, 3, 1, False, False),     # transition1[0]: bare conv, 18 output channels
    (256, 36, 3, 2, True, False),      # new-branch transition (stride 2)
    (256, 256, 3, 1, False, False),    # HRFPN fpn_conv
    (128, 128, 3, 1, True, True),
]
GRIDS = [(2, 96, 160), (1, 24, 40), (3, 13, 21), (1, 8, 16), (2, 5, 3)]


@pytest.mark.parametrize('B,H,W', GRIDS)
@pytest.mark.parametrize('cin,cout,k,stride,relu,resid', LAYERS)
def test_convgemm(built_lib, cin, cout, k, stride, relu, resid, B, H, W):
    from hrfuser_b200 import ops
    torch.manual_seed(cin + cout + k + H)
    conv = nn.Conv2d(cin, cout, k, stride, k // 2, bias=(cout == 18))
    bn = nn.BatchNorm2d(cout)
    randomize_parameters(nn.Sequential(conv, bn), cin + cout)
    bn.eval()
    assert ops.convgemm_supported(cin, cout, k, stride)
    blob = ops.pack_convgemm(conv, bn, bn.eps).cuda()
    x = (torch.randn(B, H, W, cin) * 1.5).to(torch.bfloat16)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    r = (torch.randn(B, Ho, Wo, cout)).to(torch.bfloat16) if resid else None
    got = ops.conv_gemm(x.cuda(), blob, cout, k, stride, relu, r.cuda() if resid else None)
    torch.cuda.synchronize()
    with torch.no_grad():
        y = bn(conv(x.float().permute(0, 3, 1, 2))).permute(0, 2, 3, 1)
        if resid:
            y = y + r.float()
        if relu:
            y = y.relu()
    assert tuple(got.shape) == (B, Ho, Wo, cout)
    assert_parity(got, y, 'bf16', f'convgemm {cin}->{cout} k{k} s{stride} {B}x{H}x{W}')


def test_convgemm_rejects_what_it_does_not_cover(built_lib):
    from hrfuser_b200 import ops
    assert not ops.convgemm_supported(18, 36, 3, 1)      # Cin not a multiple of 64: conv3x3_tc's job
    assert not ops.convgemm_supported(64, 512, 1, 1)
    assert not ops.convgemm_supported(64, 64, 1, 2)
    assert not ops.convgemm_supported(64, 64, 5, 1)


@pytest.mark.parametrize('cin,cout,k,stride,resid', [(256, 64, 1, 1, False), (64, 256, 1, 1, True),
                                                     (64, 64, 3, 1, False), (256, 18, 3, 1, False),
                                                     (256, 36, 3, 2, False)])
def test_convgemm_grouped(built_lib, cin, cout, k, stride, resid):
    """(1 + M) streams, one launch: three tensors of one shape, each with its own weights and
    its own ReLU flag (the camera's branch-0 transition is a bare conv, the modalities' are
    conv-BN-ReLU: hrfuser_hrformer_based.py:550-551)."""
    from hrfuser_b200 import ops
    B, H, W, n = 2, 24, 40, 3
    relus = [False, True, True]
    xs, blobs, refs, rs = [], [], [], []
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    for q in range(n):
        torch.manual_seed(10 * q + cin)
        conv = nn.Conv2d(cin, cout, k, stride, k // 2, bias=False)
        bn = nn.BatchNorm2d(cout)
        randomize_parameters(nn.Sequential(conv, bn), 3 * q + 1)
        bn.eval()
        blobs.append(ops.pack_convgemm(conv, bn, bn.eps).cuda())
        x = (torch.randn(B, H, W, cin) * 1.5).to(torch.bfloat16)
        r = torch.randn(B, Ho, Wo, cout).to(torch.bfloat16) if resid else None
        with torch.no_grad():
            y = bn(conv(x.float().permute(0, 3, 1, 2))).permute(0, 2, 3, 1)
            if resid:
                y = y + r.float()
            if relus[q]:
                y = y.relu()
        xs.append(x.cuda())
        rs.append(r.cuda() if resid else None)
        refs.append(y)
    n0 = ops.launch_count()
    got = ops.conv_gemm_grouped(xs, blobs, cout, k, stride, relus, rs if resid else None)
    torch.cuda.synchronize()
    assert ops.launch_count() - n0 == 1
    for q in range(n):
        assert_parity(got[q], refs[q], 'bf16', f'grouped convgemm problem {q} {cin}->{cout}')


@pytest.mark.parametrize('c1,c2,cout,B,H,W,n', [(64, 64, 256, 2, 96, 160, 3), (64, 64, 256, 1, 13, 21, 1),
                                                (64, 128, 128, 2, 24, 40, 2), (128, 64, 64, 1, 8, 16, 4)])
def test_convgemm_cat(built_lib, c1, c2, cout, B, H, W, n):
    """relu(bn3(conv3(y)) + bn_d(downsample(x))) (reference resnet.py:263-302) as ONE GEMM over
    the concatenated K = [y | x] (hrf_convgemm_grouped_cat_fwd), n streams in one launch, against
    the two convolutions in fp32 on the same bf16-rounded operands."""
    from hrfuser_b200 import ops
    ys, xs, blobs, refs = [], [], [], []
    relus = [q != 1 for q in range(n)]
    for q in range(n):
        torch.manual_seed(7 * q + c1 + cout + H)
        conv3, bn3 = nn.Conv2d(c1, cout, 1, bias=False), nn.BatchNorm2d(cout)
        down, bnd = nn.Conv2d(c2, cout, 1, bias=False), nn.BatchNorm2d(cout)
        randomize_parameters(nn.Sequential(conv3, bn3, down, bnd), 5 * q + 2)
        bn3.eval(), bnd.eval()
        with torch.no_grad():
            s3 = bn3.weight / torch.sqrt(bn3.running_var + bn3.eps)
            sd = bnd.weight / torch.sqrt(bnd.running_var + bnd.eps)
            both = nn.Conv2d(c1 + c2, cout, 1, bias=True)
            both.weight.copy_(torch.cat([conv3.weight * s3.view(-1, 1, 1, 1), down.weight * sd.view(-1, 1, 1, 1)], 1))
            both.bias.copy_(bn3.bias - bn3.running_mean * s3 + bnd.bias - bnd.running_mean * sd)
        blobs.append(ops.pack_convgemm(both, None).cuda())
        y = (torch.randn(B, H, W, c1) * 1.5).to(torch.bfloat16)
        x = (torch.randn(B, H, W, c2) * 1.5).to(torch.bfloat16)
        with torch.no_grad():
            ref = (bn3(conv3(y.float().permute(0, 3, 1, 2))) + bnd(down(x.float().permute(0, 3, 1, 2)))).permute(0, 2, 3, 1)
        refs.append(ref.relu() if relus[q] else ref)
        ys.append(y.cuda())
        xs.append(x.cuda())
    n0 = ops.launch_count()
    got = ops.conv_gemm_cat_grouped(ys, xs, blobs, cout, relus)
    torch.cuda.synchronize()
    assert ops.launch_count() - n0 == 1
    for q in range(n):
        assert tuple(got[q].shape) == (B, H, W, cout)
        assert_parity(got[q], refs[q], 'bf16', f'cat convgemm problem {q} {c1}+{c2}->{cout}')
    with pytest.raises(RuntimeError):                    # 32 + c2 channels: not chunks of 64
        ops.conv_gemm_cat_grouped([ys[0][..., :32].contiguous()], [xs[0]], blobs[:1], cout, True)
