"""GPU parity of every C-ABI op against the oracle (oracle/hrfuser_oracle.py)."""
import pytest
import torch

from helpers import (EDGE, NUS_B_FUSED, NUS_B_WIDE, NUS_T, STF_T, assert_parity, make_block,
                     make_exchange, tokens)
from oracle import hrfuser_oracle as O

pytestmark = pytest.mark.gpu
DT = {'fp32': torch.float32, 'bf16': torch.bfloat16}


def _engine_stub():
    """pack with the real packers through a throw-away engine-like holder"""
    from hrfuser_b200.engine import BackboneEngine
    e = BackboneEngine.__new__(BackboneEngine)
    e._host_blobs, e._blob_slots = [], []
    e.device = torch.device('cuda')
    e.dtype = torch.float32
    return e


def _nlc(t):
    return t.reshape(t.shape[0], -1, t.shape[-1])


def _ref_lsa(x, sd, heads, H, W, pad_mask=False, win=7):
    t = _nlc(x)
    n = O.layer_norm(t, sd, 'blk.norm1')
    return (t + O.window_attention(n, n, sd, 'blk.attn', H, W, heads, False, Wh=win, Ww=win,
                                   with_pad_mask=pad_mask)).view_as(x)


def _ref_mwca(x, zs, sd, heads, H, W, win=7):
    cam = _nlc(x)
    acc = cam.clone()
    for k, z in enumerate(zs):
        zt = _nlc(z)
        acc = acc + zt + O.window_attention(O.layer_norm(cam, sd, f'blk.norm1.{k}'),
                                            O.layer_norm(zt, sd, f'blk.norm2.{k}'),
                                            sd, f'blk.attn.{k}', H, W, heads, True, Wh=win, Ww=win)
    return acc.view_as(x)


SHAPES = NUS_T + STF_T + NUS_B_FUSED + NUS_B_WIDE + EDGE


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
@pytest.mark.parametrize('H,W,C,heads', SHAPES)
def test_lsa(built_lib, H, W, C, heads, mode):
    from hrfuser_b200 import ops
    B = 2
    blk, sd = make_block('lsa', C, heads, seed=H + C)
    e = _engine_stub()
    packed = e._hrformer_block(blk)
    e._upload()
    x = tokens(B, H, W, C, seed=1).to(DT[mode])
    got = ops.window_attention(x.cuda(), None, [s.t for s in packed['attn']], heads)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = _ref_lsa(x.float(), sd, heads, H, W)
    assert_parity(got, ref, mode, f'lsa {H}x{W} C{C}')


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
@pytest.mark.parametrize('M', [1, 2, 3])
@pytest.mark.parametrize('H,W,C,heads', NUS_T + STF_T + NUS_B_FUSED + NUS_B_WIDE + EDGE[:4])
def test_mwca(built_lib, H, W, C, heads, M, mode):
    from hrfuser_b200 import ops
    B = 2
    blk, sd = make_block('mwca', C, heads, M=M, seed=H + C + M)
    e = _engine_stub()
    packed = e._fusion_block(blk)
    e._upload()
    x = tokens(B, H, W, C, seed=1).to(DT[mode])
    zs = [tokens(B, H, W, C, seed=2 + k).to(DT[mode]) for k in range(M)]
    if M == 3:
        zs[1].zero_()            # a dropped modality (RandomDrop zeroes whole sensors)
    got = ops.window_attention(x.cuda(), [z.cuda() for z in zs], [s.t for s in packed['attn']], heads)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = _ref_mwca(x.float(), [z.float() for z in zs], sd, heads, H, W)
    assert_parity(got, ref, mode, f'mwca {H}x{W} C{C} M{M}')


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
@pytest.mark.parametrize('win,H,W,C,heads', [
    (14, 96, 160, 18, 1), (14, 48, 80, 36, 2), (14, 24, 40, 72, 4), (14, 12, 20, 144, 8),   # configs[4]
    (14, 24, 40, 78, 2), (14, 12, 20, 312, 8),                                              # head_dim 39
    (5, 13, 9, 36, 2), (16, 20, 33, 18, 1), (3, 7, 7, 72, 4), (1, 4, 5, 18, 1)])            # 16: largest
def test_window_sizes(built_lib, win, H, W, C, heads, mode):
    """Windows other than 7 (BASELINE.json configs[4] sweeps 7 and 14; the C-ABI takes
    win^2 <= 256): LSA and 2-modality MWCA against the oracle."""
    from hrfuser_b200 import ops
    B = 2
    e = _engine_stub()
    blk, sd = make_block('lsa', C, heads, seed=win + C, win=win)
    fblk, fsd = make_block('mwca', C, heads, M=2, seed=win + C + 1, win=win)
    p_lsa, p_mwca = e._hrformer_block(blk), e._fusion_block(fblk)
    e._upload()
    assert p_lsa['win'] == win and p_mwca['win'] == win
    x = tokens(B, H, W, C, seed=1).to(DT[mode])
    zs = [tokens(B, H, W, C, seed=2 + k).to(DT[mode]) for k in range(2)]
    got = ops.window_attention(x.cuda(), None, [s.t for s in p_lsa['attn']], heads, win=win)
    got2 = ops.window_attention(x.cuda(), [z.cuda() for z in zs], [s.t for s in p_mwca['attn']],
                                heads, win=win)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = _ref_lsa(x.float(), sd, heads, H, W, win=win)
        ref2 = _ref_mwca(x.float(), [z.float() for z in zs], fsd, heads, H, W, win=win)
    assert_parity(got, ref, mode, f'lsa win{win} {H}x{W} C{C}')
    assert_parity(got2, ref2, mode, f'mwca win{win} {H}x{W} C{C}')


def test_lsa_pad_mask(built_lib):
    from hrfuser_b200 import ops
    for (H, W) in [(12, 20), (14, 20), (5, 3)]:     # (14,20): pad_h == 0 -> mask inactive
        blk, sd = make_block('lsa', 36, 2, seed=3, with_pad_mask=True)
        e = _engine_stub()
        packed = e._hrformer_block(blk)
        e._upload()
        x = tokens(2, H, W, 36, seed=4)
        got = ops.window_attention(x.cuda(), None, [s.t for s in packed['attn']], 2, with_pad_mask=True)
        with torch.no_grad():
            ref = _ref_lsa(x, sd, 2, H, W, pad_mask=True)
        assert_parity(got, ref, 'fp32', f'lsa pad-mask {H}x{W}')


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
@pytest.mark.parametrize('H,W,C,heads', SHAPES)
def test_mixffn(built_lib, H, W, C, heads, mode):
    from hrfuser_b200 import ops
    B = 2
    blk, sd = make_block('lsa', C, heads, seed=H + C + 7)
    e = _engine_stub()
    f = e._ffn(blk.norm2, blk.ffn)
    e._upload()
    x = tokens(B, H, W, C, seed=5).to(DT[mode])
    got = ops.mixffn(x.cuda(), f['blob'].t, f['hidden'], f['eps'])
    torch.cuda.synchronize()
    with torch.no_grad():
        t = _nlc(x.float())
        ref = (t + O.cross_ffn(O.layer_norm(t, sd, 'blk.norm2'), sd, 'blk.ffn', H, W)).view_as(x)
    assert_parity(got, ref, mode, f'ffn {H}x{W} C{C}')


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
@pytest.mark.parametrize('H,W,channels,heads', [
    (96, 160, (18, 36), (1, 2)), (96, 160, (18, 36, 72), (1, 2, 4)),
    (96, 160, (18, 36, 72, 144), (1, 2, 4, 8)), (96, 312, (18, 36, 72, 144), (1, 2, 4, 8)),
    (32, 32, (78, 156), (2, 4)), (8, 24, (18, 36, 72), (1, 2, 4))])
def test_exchange(built_lib, H, W, channels, heads, mode):
    """HRModule fuse step: every output branch, incl. the fp32 NCHW copy."""
    from hrfuser_b200 import ops
    B = 2
    mod, sd = make_exchange(channels, heads, seed=len(channels))
    e = _engine_stub()
    stage = e._stage([mod])
    e._upload()
    xs = [tokens(B, H >> i, W >> i, c, seed=10 + i).to(DT[mode]) for i, c in enumerate(channels)]
    (branches, rows), = stage
    e.ops = ops
    outs, nchw = [], []
    ys = [x.cuda() for x in xs]
    for i, (ups, downs) in enumerate(rows):
        up_t = [ops.pointwise(ys[j], blob.t, cout) for j, blob, cout in ups]
        same_t = []
        for j, chain in downs:
            t = ys[j]
            for blob, cout, relu in chain:
                t = ops.dw_down(t, blob.t, cout, relu)
            same_t.append(t)
        o, n = ops.fuse_sum(ys[i], up_t, same_t, relu=True, nchw_out=True)
        outs.append(o)
        nchw.append(n)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = O.hr_exchange([x.float().permute(0, 3, 1, 2) for x in xs], sd, 'm')
    for i, r in enumerate(ref):
        assert_parity(outs[i].permute(0, 3, 1, 2), r, mode, f'exchange out{i}')
        assert nchw[i].is_contiguous() and nchw[i].dtype == torch.float32
        assert torch.equal(nchw[i].cpu(), outs[i].float().permute(0, 3, 1, 2).cpu())


def test_layout_roundtrip(built_lib):
    from hrfuser_b200 import ops
    for (B, C, H, W) in [(2, 18, 12, 20), (1, 3, 33, 65), (2, 144, 5, 7)]:
        x = torch.randn(B, C, H, W, device='cuda')
        t = ops.nchw_to_nhwc(x)
        assert torch.equal(t, x.permute(0, 2, 3, 1).contiguous())
        assert torch.equal(ops.nhwc_to_nchw(t), x)
        tb = ops.nchw_to_nhwc(x, torch.bfloat16)
        assert torch.equal(tb, x.permute(0, 2, 3, 1).contiguous().bfloat16())
        assert torch.equal(ops.nhwc_to_nchw(tb, torch.float32), x.bfloat16().float())


def test_error_paths(built_lib):
    """bad descriptors come back as error codes with a message, never a crash"""
    from hrfuser_b200 import _lib, ops
    x = torch.zeros(1, 7, 7, 18, device='cuda')
    blob = torch.zeros(10, device='cuda')
    with pytest.raises(_lib.HrfError, match='divisible'):
        ops.window_attention(x, None, [blob], heads=4)
    with pytest.raises(_lib.HrfError, match='alias'):
        ops.mixffn(x, blob, 72, out=x)
    with pytest.raises(_lib.HrfError, match='stride'):
        ops.conv3x3(x, blob, 18, stride=3)
    with pytest.raises(_lib.HrfError, match='even'):
        ops.conv3x3(torch.zeros(1, 7, 7, 17, device='cuda'), blob, 18)


def test_programmatic_dependent_launch(built_lib):
    """hrf_set_pdl(1): a chain lsa -> mixffn -> lsa launched with programmatic stream
    serialization (each kernel's prologue overlaps its predecessor's tail) gives the same bits"""
    from hrfuser_b200 import ops
    blk, sd = make_block('lsa', 18, 1)
    e = _engine_stub()
    pk = e._hrformer_block(blk)
    e._upload()
    blobs, f = [s.t for s in pk['attn']], pk['ffn']
    x = tokens(2, 33, 47, 18, seed=5).to(torch.bfloat16).cuda()

    def chain():
        y = x
        for _ in range(3):
            y = ops.window_attention(y, None, blobs, 1)
            y = ops.mixffn(y, f['blob'].t, f['hidden'], f['eps'])
        torch.cuda.synchronize()
        return y
    was = ops.set_pdl(False)
    try:
        ref = chain()
        assert ops.set_pdl(True) is False
        got = chain()
    finally:
        ops.set_pdl(was)
    assert torch.equal(ref, got)


@pytest.mark.parametrize('dt', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('C,with_res,relu', [(64, False, True), (256, True, True), (18, False, False),
                                             (36, True, True), (3, False, True)])
def test_bias_act(built_lib, C, with_res, relu, dt):
    from hrfuser_b200 import ops
    g = torch.Generator().manual_seed(C)
    y = torch.randn(2, 9, 13, C, generator=g).to(dt).cuda()
    r = torch.randn(2, 9, 13, C, generator=g).to(dt).cuda() if with_res else None
    b = torch.randn(C, generator=g).cuda()
    ref = (y.float() + r.float() if with_res else y.float()) + b      # kernel's order of adds
    ref = (ref.relu() if relu else ref).to(dt)
    ops.bias_act_(y, b, r, relu)
    torch.cuda.synchronize()
    assert torch.equal(y, ref)


@pytest.mark.parametrize('cin,H,W', [(3, 64, 96), (2, 34, 50), (1, 33, 47), (3, 384, 640)])
def test_stem_conv_tc(built_lib, cin, H, W):
    """tcgen05 stem conv (fp32 NCHW in, bf16 NHWC out) vs conv2d + eval BN + ReLU"""
    import torch.nn as nn
    import torch.nn.functional as F
    from hrfuser_b200 import ops
    from hrfuser_b200.utils import randomize_parameters
    conv, bn = nn.Conv2d(cin, 64, 3, 2, 1, bias=False), nn.BatchNorm2d(64)
    randomize_parameters(nn.Sequential(conv, bn), cin)
    bn.eval()
    x = torch.randn(2, cin, H, W, generator=torch.Generator().manual_seed(H))
    with torch.no_grad():
        ref = F.relu(bn(conv(x))).permute(0, 2, 3, 1)
    blob = ops.pack_stem(conv, bn, bn.eps).cuda()
    got = ops.stem_conv(x.cuda(), blob, 64)
    torch.cuda.synchronize()
    assert got.dtype == torch.bfloat16 and tuple(got.shape) == tuple(ref.shape)
    assert_parity(got, ref, 'bf16', f'stem conv cin={cin}')


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
@pytest.mark.parametrize('cin,cout,stride,H,W,relu', [
    (18, 36, 2, 96, 160, True), (36, 72, 2, 48, 80, True), (72, 144, 2, 24, 40, True),
    (18, 36, 2, 33, 47, True), (36, 72, 2, 7, 5, False), (18, 18, 1, 14, 21, True), (72, 144, 2, 1, 1, True),
    (18, 20, 2, 9, 11, True), (78, 156, 2, 12, 20, True), (14, 32, 1, 6, 5, False)])
def test_conv3x3(built_lib, cin, cout, stride, H, W, relu, mode):
    """transition conv (3x3, pad 1, stride 1|2) + eval BN (+ReLU) on tokens vs conv2d"""
    import torch.nn as nn
    import torch.nn.functional as F
    from hrfuser_b200 import ops
    from hrfuser_b200.utils import randomize_parameters
    conv, bn = nn.Conv2d(cin, cout, 3, stride, 1, bias=False), nn.BatchNorm2d(cout)
    randomize_parameters(nn.Sequential(conv, bn), cin + cout)
    bn.eval()
    B = 2
    x = tokens(B, H, W, cin, seed=H + W).to(DT[mode])
    with torch.no_grad():
        ref = bn(conv(x.float().permute(0, 3, 1, 2)))
        ref = (F.relu(ref) if relu else ref).permute(0, 2, 3, 1)
    blob = ops.pack_conv3x3(conv, bn, bn.eps).cuda()
    got = ops.conv3x3(x.cuda(), blob, cout, stride, relu)
    torch.cuda.synchronize()
    assert got.dtype == DT[mode] and tuple(got.shape) == tuple(ref.shape)
    assert_parity(got, ref, mode, f'conv3x3 {cin}->{cout} s{stride} {H}x{W}')


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
@pytest.mark.parametrize('H,W,C,heads', [NUS_T[0], STF_T[0], NUS_T[1], EDGE[2], (13, 9, 18, 1)])
@pytest.mark.parametrize('n', [1, 2, 3, 4])
def test_grouped_lsa_and_mixffn(built_lib, H, W, C, heads, n, mode):
    """hrf_window_attn_grouped_fwd / hrf_mixffn_grouped_fwd: the camera's branch 0 and the modality
    streams in one launch -- n tensors of one shape, each with its own weights; every problem must
    equal its single-tensor call's reference (one launch for C = 18 in bf16 mode)."""
    from hrfuser_b200 import ops
    B = 3
    xs, attn_blobs, ffn_packs, sds = [], [], [], []
    e = _engine_stub()
    for q in range(n):
        blk, sd = make_block('lsa', C, heads, seed=10 * q + H + C)
        packed = e._hrformer_block(blk)
        attn_blobs.append(packed['attn'][0])
        ffn_packs.append(packed['ffn'])
        sds.append(sd)
        xs.append(tokens(B, H, W, C, seed=q + 1).to(DT[mode]))
    e._upload()
    xc = [x.cuda() for x in xs]
    n0 = ops.launch_count()
    got_a = ops.window_attention_grouped(xc, [b.t for b in attn_blobs], heads)
    n1 = ops.launch_count()
    got_f = ops.mixffn_grouped(xc, [f['blob'].t for f in ffn_packs], ffn_packs[0]['hidden'], ffn_packs[0]['eps'])
    n2 = ops.launch_count()
    torch.cuda.synchronize()
    if mode == 'bf16' and C == 18:
        assert n1 - n0 == 1                            # one launch for the whole group
        if (W * C * 2) % 16 == 0:                      # (MixFFN: rows the TMA unit can address)
            assert n2 - n1 == 1
    for q in range(n):
        with torch.no_grad():
            ref_a = _ref_lsa(xs[q].float(), sds[q], heads, H, W)
            t = xs[q].float().view(B, H * W, C)
            ref_f = (t + O.cross_ffn(O.layer_norm(t, sds[q], 'blk.norm2'), sds[q], 'blk.ffn', H, W)).view_as(xs[q])
        assert_parity(got_a[q], ref_a, mode, f'grouped lsa {q}/{n} {H}x{W} C{C}')
        assert_parity(got_f[q], ref_f, mode, f'grouped mixffn {q}/{n} {H}x{W} C{C}')
