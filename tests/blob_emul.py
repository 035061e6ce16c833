"""CPU emulation of the device ops *from the packed blobs* (test infrastructure).

Mirrors the blob layouts of hrfuser_b200/csrc/{window_attn,mixffn,hrfuse}.cuh in
Python and evaluates each op with torch CPU ops.  Plugged into
`BackboneEngine(device_ops=...)` it checks, without a GPU, that
  * the C packers (hrf_*_pack: BN folding, transposes, head padding, q scaling),
  * the engine's wiring (quirks included)
reproduce the oracle.  The kernels' device code itself is checked on the GPU.
"""
import math

import torch
import torch.nn.functional as F


def _ru(a, b):
    return (a + b - 1) // b * b


class AttnLayout:
    def __init__(self, C, heads, win):
        self.C, self.heads = C, heads
        self.hd = C // heads
        self.hdp = _ru(self.hd, 4)
        self.Cp = _ru(C, 4)
        self.KO = heads * self.hdp
        self.S = win * win
        self.T = (2 * win - 1) ** 2
        c4 = _ru(C, 4)
        o = 0
        names = {}
        for n in ('lnq_w', 'lnq_b', 'lnkv_w', 'lnkv_b'):
            names[n] = o
            o += c4
        for w, b in (('wq', 'bq'), ('wk', 'bk'), ('wv', 'bv')):
            names[w] = o
            o = _ru(o + self.Cp * C, 4)
            names[b] = o
            o += c4
        names['wo'] = o
        o = _ru(o + self.KO * C, 4)
        names['bo'] = o
        o += c4
        names['rpb'] = o
        o += _ru(heads * self.T, 4)
        # tensor-core sections (bf16 operand tiles; checked on the GPU, skipped here)
        HDP, KC = _ru(self.hd, 16), _ru(C, 16)
        NQ = heads * HDP
        names['tc_bias'] = o
        o += 3 * NQ + KC
        for w in ('tc_wq', 'tc_wk', 'tc_wv'):
            names[w] = o
            o += NQ * KC // 2
        names['tc_wo'] = o
        o += KC * NQ // 2
        self.tc = dict(HDP=HDP, KC=KC, NQ=NQ, NOUT=KC)
        if C == 18 and heads == 1 and win == 7:        # third-generation sections (window_attn_v3.cuh)
            o = _ru(o, 4)
            names['v3'] = o
            o += (4784 + 5808) // 4
        self.total = o
        self.o = names


def window_attention(x, kv, blobs, heads, win=7, with_pad_mask=False, eps=1e-6, out=None):
    B, H, W, C = x.shape
    L = AttnLayout(C, heads, win)
    kv = list(kv or [])
    cross = len(kv) > 0
    nWh, nWw = -(-H // win), -(-W // win)
    ph, pw = nWh * win - H, nWw * win - W
    pt, pl = ph // 2, pw // 2
    xf = x.float()
    acc = xf.clone()
    for k in range(max(1, len(kv))):
        blob = blobs[k].float().cpu()
        assert blob.numel() == L.total
        g = lambda n, sz: blob[L.o[n]:L.o[n] + sz]
        zf = kv[k].float() if cross else xf
        xn = F.layer_norm(xf, (C,), g('lnq_w', C), g('lnq_b', C), eps)
        zn = F.layer_norm(zf, (C,), g('lnkv_w', C), g('lnkv_b', C), eps) if cross else xn

        def part(t):
            t = F.pad(t, (0, 0, pl, pw - pl, pt, ph - pt))
            return t.view(B, nWh, win, nWw, win, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, L.S, C)
        wmat = lambda n: blob[L.o[n]:L.o[n] + L.Cp * C].view(L.Cp, C)[:C]      # k-major
        q = part(xn) @ wmat('wq') + g('bq', C)          # already scaled by hd^-0.5
        kk = part(zn) @ wmat('wk') + g('bk', C)
        v = part(zn) @ wmat('wv') + g('bv', C)
        sp = lambda t: t.view(-1, L.S, heads, L.hd).transpose(1, 2)
        logits = sp(q) @ sp(kk).transpose(-1, -2)
        table = blob[L.o['rpb']:L.o['rpb'] + heads * L.T].view(heads, L.T)
        s = torch.arange(L.S)
        hh, ww = s // win, s % win
        idx = (hh[:, None] - hh[None] + win - 1) * (2 * win - 1) + (ww[:, None] - ww[None] + win - 1)
        logits = logits + table[:, idx][None]
        if with_pad_mask and ph > 0 and pw > 0:
            m = F.pad(torch.zeros(1, H, W, 1), (0, 0, pl, pw - pl, pt, ph - pt), value=-math.inf)
            m = m.view(1, nWh, win, nWw, win, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, L.S)
            logits = logits + m.repeat(B, 1)[:, None, None, :]
        o = logits.softmax(-1) @ sp(v)                                   # (nW, h, S, hd)
        # head-padded O layout feeding WoT [KO][C]
        opad = torch.zeros(o.shape[0], L.S, L.KO)
        for h in range(heads):
            opad[:, :, h * L.hdp:h * L.hdp + L.hd] = o[:, h]
        wo = blob[L.o['wo']:L.o['wo'] + L.KO * C].view(L.KO, C)
        y = opad @ wo + g('bo', C)
        y = y.view(B, nWh, nWw, win, win, C).permute(0, 1, 3, 2, 4, 5).reshape(B, nWh * win, nWw * win, C)
        y = y[:, pt:pt + H, pl:pl + W]
        acc = acc + y + (zf if cross else 0)
    return acc.to(x.dtype)


class FfnLayout:
    def __init__(self, C, hidden):
        self.C, self.hidden, self.Cp = C, hidden, _ru(C, 4)
        c4 = _ru(C, 4)
        o = 0
        self.ln_w = o; o += c4
        self.ln_b = o; o += c4
        self.w1 = o; o += self.Cp * hidden
        self.b1 = o; o += hidden
        self.wd = o; o += 9 * hidden
        self.bd = o; o += hidden
        self.w2 = o; o = _ru(o + hidden * C, 4)
        self.b2 = o; o += c4
        # tensor-core sections (bf16 tiles; exercised on the GPU)
        KC = _ru(C, 16)
        nch = hidden // 72 if hidden % 72 == 0 else hidden // 78 if hidden % 78 == 0 else 0
        self.tc_f32 = o
        o = _ru(o + nch * 880 + (KC if nch else 0), 4)
        self.tc_w1 = o; o += nch * 80 * KC // 2
        self.tc_w2 = o; o += nch * KC * 80 // 2
        v2 = nch if hidden % 72 == 0 else 0          # second-generation kernel sections
        self.v2_w1 = o; o += v2 * 80 * KC // 2
        self.v2_w2 = o; o += v2 * KC * 80 // 2
        self.v2_cv = o; o += v2 * 400
        self.tc = dict(KC=KC, NOUT=KC, nchunk=nch)
        self.total = o


def _gelu(x):
    return 0.5 * x * (1 + torch.erf(x / math.sqrt(2.0)))


def mixffn(x, blob, hidden, eps=1e-6, out=None):
    B, H, W, C = x.shape
    L = FfnLayout(C, hidden)
    blob = blob.float().cpu()
    assert blob.numel() == L.total
    xf = x.float()
    xn = F.layer_norm(xf, (C,), blob[L.ln_w:L.ln_w + C], blob[L.ln_b:L.ln_b + C], eps)
    w1 = blob[L.w1:L.w1 + L.Cp * hidden].view(L.Cp, hidden)[:C]
    h1 = _gelu(xn @ w1 + blob[L.b1:L.b1 + hidden])
    wd = blob[L.wd:L.wd + 9 * hidden].view(3, 3, hidden).permute(2, 0, 1)[:, None]
    h2 = F.conv2d(h1.permute(0, 3, 1, 2), wd, blob[L.bd:L.bd + hidden], 1, 1, 1, hidden)
    h2 = _gelu(h2).permute(0, 2, 3, 1)
    w2 = blob[L.w2:L.w2 + hidden * C].view(hidden, C)
    y = _gelu(h2 @ w2 + blob[L.b2:L.b2 + C])
    return (xf + y).to(x.dtype)


def _pw(x, blob, cout, relu):
    Cin = x.shape[-1]
    Kp = _ru(Cin, 4)
    w = blob[:Kp * cout].view(Kp, cout)[:Cin]
    ob = _ru(Kp * cout, 4)
    y = x @ w + blob[ob:ob + cout]
    return y.relu() if relu else y


def pointwise(x, blob, cout, relu=False):
    blob = blob.float().cpu()
    assert blob.numel() == _ru(_ru(x.shape[-1], 4) * cout, 4) + _ru(cout, 4)
    return _pw(x.float(), blob, cout, relu).to(x.dtype)


def conv3x3(x, blob, cout, stride=1, relu=False):
    """blob = PwLayout(9*Cin, cout): Wt [Kp][cout], k = tap*Cin + c; then the bias."""
    blob = blob.float().cpu()
    Cin = x.shape[-1]
    K = 9 * Cin
    Kp = _ru(K, 4)
    base = _ru(Kp * cout, 4) + _ru(cout, 4)
    # bf16 tiles of the tensor-core kernel (ConvTcCfg in csrc/conv3x3_tc.cuh)
    nsplit = 2 if (Cin, cout) == (72, 144) else 1
    nout, kc = _ru(cout // nsplit, 16), _ru(Cin + 1, 16)
    ok = (Cin == 18 and nout in (32, 48, 80, 144)) or (Cin, cout) in ((36, 72), (72, 144))
    tc = nsplit * 9 * nout * kc // 2 if ok else 0
    assert blob.numel() == (_ru(base, 4) + tc if tc else base)
    w = blob[:Kp * cout].view(Kp, cout)[:K].view(3, 3, Cin, cout).permute(3, 2, 0, 1)
    ob = _ru(Kp * cout, 4)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w, blob[ob:ob + cout], stride, 1)
    y = y.relu() if relu else y
    return y.permute(0, 2, 3, 1).contiguous().to(x.dtype)


def dw_down(x, blob, cout, relu=False):
    blob = blob.float().cpu()
    Cin = x.shape[-1]
    o_bd = _ru(9 * Cin, 4)
    o_pw = o_bd + _ru(Cin, 4)
    wd = blob[:9 * Cin].view(3, 3, Cin).permute(2, 0, 1)[:, None]
    t = F.conv2d(x.float().permute(0, 3, 1, 2), wd, blob[o_bd:o_bd + Cin], 2, 1, 1, Cin)
    return _pw(t.permute(0, 2, 3, 1), blob[o_pw:], cout, relu).to(x.dtype)


def fuse_sum(x, ups=(), sames=(), relu=True, nchw_out=False):
    B, H, W, C = x.shape
    v = x.float()
    for s in sames:
        v = v + s.float()
    for u in ups:
        v = v + F.interpolate(u.float().permute(0, 3, 1, 2), size=(H, W), mode='bilinear',
                              align_corners=False).permute(0, 2, 3, 1)
    if relu:
        v = v.relu()
    out = v.to(x.dtype).contiguous()
    if nchw_out:
        return out, out.float().permute(0, 3, 1, 2).contiguous()
    return out
