"""The third-generation attention kernel (csrc/window_attn_v3.cuh, C = 18 / one head) runs on
operand tiles in which hrf_attn_pack has folded the LayerNorm affine, the softmax scale, log2 e
and the output projection.  This test decodes those tiles from the blob and evaluates the
kernel's arithmetic in torch on the CPU -- x^ = (x - mean) rstd with the two constant columns,
one packed projection, base-2 softmax numerators, denominator from the constant-1 column of V' --
against the oracle's window attention (reference hrformer.py:96-131,184-236 /
hrfuser_hrformer_based.py:106-151,189-248) on a grid that needs padding."""
import math

import pytest
import torch

from oracle import hrfuser_oracle as O

C, S = 18, 49
SEC_SELF, W_Q_B, W_KV_B, W_SELF_B = 4784, 2048, 3072, 4096


def _tile(raw_bytes, rows, cols=32):
    t = raw_bytes.view(torch.bfloat16).float()
    return t.view(cols // 8, rows, 8).permute(1, 0, 2).reshape(rows, cols)


def _sections(blob):
    from blob_emul import AttnLayout
    L = AttnLayout(C, 1, 7)
    b = blob[L.o['v3']:L.o['v3'] + (4784 + 5808) // 4].view(torch.uint8)
    w_self = _tile(b[:W_SELF_B], 64)
    t_self = b[W_SELF_B:SEC_SELF].view(torch.float32)[:169]
    w_q = _tile(b[SEC_SELF:SEC_SELF + W_Q_B], 32)
    w_kv = _tile(b[SEC_SELF + W_Q_B:SEC_SELF + W_Q_B + W_KV_B], 48)
    t_cross = b[SEC_SELF + W_Q_B + W_KV_B:].view(torch.float32)[:169]
    return w_self, t_self, w_q, w_kv, t_cross


def _xhat(x, valid):
    """(nW, S, C) raw window tokens -> (nW, S, 32): normalised, column 18 = 1, column 19 = real"""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    n = (x - mu) * torch.rsqrt(var + 1e-6) * valid[..., None]
    out = torch.zeros(*x.shape[:-1], 32)
    out[..., :C] = n
    out[..., 18] = 1.0
    out[..., 19] = valid.float()
    return out.bfloat16().float()


def _v3_emulate(xw, zw, valid, wq_rows, wk_rows, wv_rows, table, mask):
    xh, zh = _xhat(xw, valid), _xhat(zw, valid)
    q = (xh @ wq_rows.t()).bfloat16().float()
    k = (zh @ wk_rows.t()).bfloat16().float()
    v = (zh @ wv_rows.t()).bfloat16().float()                       # (nW, S, 19), column 18 == 1
    rpi = torch.from_numpy(O.relative_position_index(7, 7)).reshape(-1).long()
    s = q @ k.transpose(-1, -2) + table[rpi].reshape(S, S)
    if mask:
        s = s + torch.where(valid, 0.0, -math.inf)[:, None, :]
    p = torch.exp2(s - s.amax(-1, keepdim=True)).bfloat16().float()
    y = p @ v
    return y[..., :C] / y[..., 18:19]


@pytest.mark.parametrize('cross', [False, True])
@pytest.mark.parametrize('mask', [False, True])
def test_v3_tiles_reproduce_window_attention(built_lib, cross, mask):
    from hrfuser_b200 import ops
    g = torch.Generator().manual_seed(11)
    r = lambda *s: torch.randn(*s, generator=g)
    H, W, B = 10, 16, 2                                              # pads on both axes
    lnq, lnk = (1 + 0.3 * r(C), 0.3 * r(C)), (1 + 0.3 * r(C), 0.3 * r(C))
    if not cross:
        lnk = lnq
    wq, wk, wv, wo = (0.3 * r(C, C) for _ in range(4))
    bq, bk, bv, bo = (0.2 * r(C) for _ in range(4))
    table = 0.5 * r(169, 1)
    blob = ops.pack_attn(C, 1, 7, lnq, lnk, wq, bq, wk, bk, wv, bv, wo, bo, table)
    w_self, t_self, w_q, w_kv, t_cross = _sections(blob)
    assert torch.equal(t_self, t_cross)
    assert torch.equal(w_self[:18], w_q[:18]) and torch.equal(w_self[18:55], w_kv[:37])
    assert w_self[55:].abs().sum() == 0 and w_q[18:].abs().sum() == 0 and w_kv[37:].abs().sum() == 0
    assert w_self[:, 20:].abs().sum() == 0
    ones_row = torch.zeros(32)
    ones_row[18] = 1.0
    assert torch.equal(w_self[54], ones_row)                         # V' column 18: the constant 1

    x = r(B, H * W, C).bfloat16().float()
    z = r(B, H * W, C).bfloat16().float() if cross else x
    # oracle: LN (with affine) -> window attention -> (B, N, C)
    sd = {'a.attn.relative_position_bias_table': table, 'a.attn.out_proj.weight': wo, 'a.attn.out_proj.bias': bo}
    if cross:
        sd.update({'a.attn.q_proj.weight': wq, 'a.attn.q_proj.bias': bq, 'a.attn.k_proj.weight': wk,
                   'a.attn.k_proj.bias': bk, 'a.attn.v_proj.weight': wv, 'a.attn.v_proj.bias': bv})
    else:
        sd.update({'a.attn.qkv.weight': torch.cat([wq, wk, wv]), 'a.attn.qkv.bias': torch.cat([bq, bk, bv])})
    ln = lambda t, p: torch.nn.functional.layer_norm(t, (C,), p[0], p[1], 1e-6)
    ref = O.window_attention(ln(x, lnq), ln(z, lnk), sd, 'a', H, W, 1, cross, with_pad_mask=mask)

    gmap = torch.from_numpy(O.window_gather_map(H, W, 7, 7)).long()
    valid = gmap >= 0
    idx = gmap.clamp_min(0)
    win = lambda t: t[:, idx.reshape(-1)].reshape(B, *gmap.shape, C)
    for b in range(B):
        y = _v3_emulate(win(x)[b], win(z)[b], valid, w_self[:18], w_self[18:36], w_self[36:55], t_self, mask)
        out = torch.zeros(H * W, C)
        out[gmap[valid]] = y[valid]
        err = (out - ref[b]).norm() / ref[b].norm()
        assert err < 1.5e-2, err                                    # bf16 operands, fp32 accumulation
