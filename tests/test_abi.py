"""The C-ABI library builds, loads on a CPU-only host, exports every symbol the
header declares, and its host-side entry points (packers, validation) behave."""
import ctypes as C
import os
import re

import pytest
import torch

from hrfuser_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'hrfuser_b200.h')


def _declared():
    txt = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r'\b(hrf_[a-z0-9_]+)\s*\(', txt)))


def test_header_symbols_exported_and_bound(built_lib):
    names = _declared()
    assert len(names) >= 19
    for n in names:
        assert hasattr(built_lib, n), f'{n} declared in the header but not exported'
        assert n in _lib.SIGNATURES, f'{n} has no ctypes signature'
    assert sorted(_lib.SIGNATURES) == names
    assert built_lib.hrf_abi_version() == _lib.ABI_VERSION
    ver = int(re.search(r'#define HRF_ABI_VERSION (\d+)', open(HEADER).read()).group(1))
    assert ver == _lib.ABI_VERSION


def test_struct_layouts_match_header():
    # plain int32/float PODs: sizes as the C compiler lays them out
    assert C.sizeof(_lib.AttnDesc) == 40
    assert C.sizeof(_lib.FfnDesc) == 28
    assert C.sizeof(_lib.PwDesc) == 28
    assert C.sizeof(_lib.FuseDesc) == 4 * (6 + 2 * _lib.MAX_FUSE_TERMS + 2)


def test_validation_errors_are_codes_with_messages(built_lib):
    d = _lib.AttnDesc(1, 7, 7, 18, 4, 7, 0, 0, 0, 1e-6)          # 18 % 4 != 0
    buf = (C.c_float * 8)()
    rc = built_lib.hrf_attn_pack(C.byref(d), *([buf] * 13), buf)
    assert rc == -1 and b'divisible' in built_lib.hrf_last_error()
    d = _lib.AttnDesc(1, 7, 7, 624, 16, 7, 0, 0, 0, 1e-6)        # wider than the fused kernels:
    assert built_lib.hrf_attn_workspace_bytes(C.byref(d)) > 0    # generic path, needs scratch
    assert built_lib.hrf_attn_blob_floats(None) == 0
    f = _lib.FfnDesc(1, 8, 8, 18, 70, 0, 1e-6)                   # hidden not a multiple of 4
    assert built_lib.hrf_mixffn_fwd(C.byref(f), None, None, None, None, 0, None) == -2
    f = _lib.FfnDesc(1, 8, 8, 18, 72, 0, 1e-6)
    assert built_lib.hrf_mixffn_fwd(C.byref(f), None, None, None, None, 0, None) == -1   # null pointers


def test_attn_packer_layout(built_lib):
    """q is pre-scaled, weights are k-major, out_proj rows follow the head-padded
    layout, the rpb table is transposed to (heads, 169)."""
    from blob_emul import AttnLayout
    from hrfuser_b200 import ops
    Cc, heads = 36, 2
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g)
    lnq, lnk = (r(Cc), r(Cc)), (r(Cc), r(Cc))
    wq, wk, wv, wo = r(Cc, Cc), r(Cc, Cc), r(Cc, Cc), r(Cc, Cc)
    bq, bk, bv, bo = r(Cc), r(Cc), r(Cc), r(Cc)
    table = r(169, heads)
    blob = ops.pack_attn(Cc, heads, 7, lnq, lnk, wq, bq, wk, bk, wv, bv, wo, bo, table)
    L = AttnLayout(Cc, heads, 7)
    assert blob.numel() == L.total
    scale = (Cc // heads) ** -0.5
    sec = lambda n, sz: blob[L.o[n]:L.o[n] + sz]
    assert torch.allclose(sec('wq', Cc * Cc).view(Cc, Cc), wq.t() * scale)
    assert torch.allclose(sec('bq', Cc), bq * scale)
    assert torch.equal(sec('wk', Cc * Cc).view(Cc, Cc), wk.t())
    assert torch.equal(sec('lnkv_b', Cc), lnk[1])
    wo_p = sec('wo', L.KO * Cc).view(L.KO, Cc)
    hd, hdp = L.hd, L.hdp
    for h in range(heads):
        assert torch.equal(wo_p[h * hdp:h * hdp + hd], wo.t()[h * hd:(h + 1) * hd])
        assert wo_p[h * hdp + hd:(h + 1) * hdp].abs().sum() == 0
    assert torch.equal(sec('rpb', heads * 169).view(heads, 169), table.t())


def test_inference_without_gpu_fails_loudly():
    import copy
    from hrfuser_b200 import HRFuserHRFormerBased, tiny_cfg
    c = copy.deepcopy(tiny_cfg(2))
    c.pop('type')
    net = HRFuserHRFormerBased(**c).eval()
    x = torch.zeros(1, 3, 32, 32)
    if not torch.cuda.is_available():
        with pytest.raises(_lib.HrfError, match='no CPU'):
            with torch.no_grad():
                net(x, [x, x])


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(_lib.HrfError, match='not built'):
        _lib.load()


def test_attn_packer_tensor_core_tiles(built_lib):
    """bf16 operand tiles of the tcgen05 kernel: chunk-major [K/8][rows][8], heads
    padded 18 -> 32, q pre-scaled, K column C = the bias (multiplies the constant-1
    column of LN(x)), out_proj K index = padded O column."""
    from blob_emul import AttnLayout
    from hrfuser_b200 import ops
    Cc, heads = 36, 2
    g = torch.Generator().manual_seed(1)
    r = lambda *s: torch.randn(*s, generator=g)
    ln = (r(Cc), r(Cc))
    wq, wk, wv, wo = r(Cc, Cc), r(Cc, Cc), r(Cc, Cc), r(Cc, Cc)
    bq, bk, bv, bo = r(Cc), r(Cc), r(Cc), r(Cc)
    blob = ops.pack_attn(Cc, heads, 7, ln, ln, wq, bq, wk, bk, wv, bv, wo, bo, r(169, heads))
    L = AttnLayout(Cc, heads, 7)
    HDP, KC, NQ, NOUT = (L.tc[k] for k in ('HDP', 'KC', 'NQ', 'NOUT'))
    hd = Cc // heads
    scale = hd ** -0.5

    def tile(name, rows, cols):
        raw = blob[L.o[name]:L.o[name] + rows * cols // 2].view(torch.bfloat16).float()
        return raw.view(cols // 8, rows, 8).permute(1, 0, 2).reshape(rows, cols)   # -> [row][col]
    pad_rows = torch.tensor([h * HDP + d for h in range(heads) for d in range(hd)])
    for name, w, b, s in (('tc_wq', wq, bq, scale), ('tc_wk', wk, bk, 1.0), ('tc_wv', wv, bv, 1.0)):
        t = tile(name, NQ, KC)
        assert torch.equal(t[pad_rows, :Cc], (w * s).bfloat16().float())
        assert KC > Cc and torch.equal(t[pad_rows, Cc], (b * s).bfloat16().float())   # bias row
        assert t[:, Cc + 1:].abs().sum() == 0
        keep = torch.zeros(NQ, dtype=torch.bool)
        keep[pad_rows] = True
        assert t[~keep].abs().sum() == 0
    to = tile('tc_wo', NOUT, NQ)
    assert torch.equal(to[:Cc][:, pad_rows], wo.bfloat16().float())
    assert to[Cc:].abs().sum() == 0
    bias = blob[L.o['tc_bias']:L.o['tc_bias'] + 3 * NQ + NOUT]
    assert torch.allclose(bias[pad_rows], bq * scale)
    assert torch.equal(bias[NQ + pad_rows], bk) and torch.equal(bias[2 * NQ + pad_rows], bv)
    assert torch.equal(bias[3 * NQ:3 * NQ + Cc], bo)


def test_conv3x3_packer(built_lib):
    """hrf_conv3x3_pack: fp32 k-major section (K = tap*Cin + c, BN folded) and, for the widths
    the tcgen05 kernel covers, nine bf16 B tiles [KC/8][NOUT][8] whose row Cin of the centre tap
    is the folded bias."""
    import torch.nn as nn
    from hrfuser_b200 import ops
    from hrfuser_b200.utils import randomize_parameters
    cin, cout = 18, 36
    conv, bn = nn.Conv2d(cin, cout, 3, 2, 1, bias=False), nn.BatchNorm2d(cout)
    randomize_parameters(nn.Sequential(conv, bn), 7)
    bn.eval()
    blob = ops.pack_conv3x3(conv, bn, bn.eps)
    sc = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    wf = (conv.weight * sc[:, None, None, None]).detach()               # (Cout, Cin, 3, 3)
    bf = (bn.bias - bn.running_mean * sc).detach()
    K, Kp = 9 * cin, (9 * cin + 3) // 4 * 4
    wt = blob[:Kp * cout].view(Kp, cout)[:K].view(9, cin, cout)          # [tap][c][n]
    assert torch.allclose(wt, wf.permute(2, 3, 1, 0).reshape(9, cin, cout), rtol=1e-6, atol=1e-7)
    ob = (Kp * cout + 3) // 4 * 4
    assert torch.allclose(blob[ob:ob + cout], bf, rtol=1e-6, atol=1e-7)
    base = ob + (cout + 3) // 4 * 4
    nout, kc = (cout + 15) // 16 * 16, (cin + 1 + 15) // 16 * 16
    assert blob.numel() == base + 9 * nout * kc // 2
    tiles = blob[base:].view(torch.bfloat16).float().view(9, kc // 8, nout, 8).permute(0, 2, 1, 3).reshape(9, nout, kc)
    ref = wf.permute(2, 3, 0, 1).reshape(9, cout, cin)
    assert torch.equal(tiles[:, :cout, :cin], ref.bfloat16().float())
    assert torch.equal(tiles[4, :cout, cin], bf.bfloat16().float())       # bias row, centre tap
    tiles[4, :cout, cin] = 0
    assert tiles[:, :, cin:].abs().sum() == 0 and tiles[:, cout:].abs().sum() == 0
