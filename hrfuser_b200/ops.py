"""Torch-tensor front end of the C-ABI ops (device memory, streams: plumbing only).

Activations are channels-last token tensors of shape (B, H, W, C), contiguous,
fp32 or bf16, on a CUDA device.  Packed weights are fp32 blobs produced by the
`pack_*` helpers (host side, via the library's own packers) and uploaded once.
"""
import contextlib
import ctypes as C

import torch

from . import _lib
from ._lib import (AttnDesc, ConvDesc, ConvGemmDesc, DwPwDesc, FfnDesc, FuseDesc, HRF_BF16, HRF_F32, HRF_U8, InputDesc,
                   PwDesc, StemDesc, check)

_DT = {torch.float32: HRF_F32, torch.bfloat16: HRF_BF16}
_F = C.POINTER(C.c_float)


def _dtype_code(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f'unsupported activation dtype {t.dtype}') from None


def _host(t):
    """detached fp32 contiguous CPU tensor (kept alive by the caller's list)"""
    return t.detach().to('cpu', torch.float32).contiguous()


def _fp(t):
    return C.cast(t.data_ptr(), _F) if t is not None else None


def _bn4(bn, keep):
    """BatchNorm module -> `const float* const[4]` (weight, bias, mean, var)."""
    if bn is None:
        return None
    ts = [_host(bn.weight), _host(bn.bias), _host(bn.running_mean), _host(bn.running_var)]
    keep.extend(ts)
    return (_F * 4)(*[_fp(t) for t in ts])


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_act(x, name='x'):
    if not (x.is_cuda and x.dim() == 4 and x.is_contiguous()):
        raise ValueError(f'{name} must be a contiguous CUDA (B,H,W,C) tensor')


class Recorder:
    """Per-call CUDA-event timing of the ops (used by bench.py for the roofline
    lines).  Events are recorded on the launching stream around each C-ABI call."""

    def __init__(self):
        self.calls = []          # (kind, meta dict, start event, end event)

    def summary(self):
        """-> {(kind, C): dict(calls, total_ms, avg_ms, bytes, flops)} after a sync"""
        out = {}
        for kind, meta, a, b in self.calls:
            g = out.setdefault((kind, meta['C']), dict(calls=0, problems=0, total_ms=0.0, bytes=0.0, flops=0.0))
            g['calls'] += 1
            g['problems'] += meta.get('problems', 1)     # grouped launches: tensors per call
            g['total_ms'] += a.elapsed_time(b)
            g['bytes'] += meta['bytes']
            g['flops'] += meta['flops']
        for g in out.values():
            g['avg_ms'] = g['total_ms'] / g['calls']
        return out


_RECORDER = None


@contextlib.contextmanager
def record():
    global _RECORDER
    prev, _RECORDER = _RECORDER, Recorder()
    try:
        yield _RECORDER
    finally:
        _RECORDER = prev


@contextlib.contextmanager
def _timed(kind, **meta):
    if _RECORDER is None:
        yield
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    yield
    b.record()
    _RECORDER.calls.append((kind, meta, a, b))


def _ptr_array(tensors):
    return (C.c_void_p * max(1, len(tensors)))(*[t.data_ptr() for t in tensors])


# ----------------------------------------------------------------------------
# packers (host)
# ----------------------------------------------------------------------------
def attn_desc(B, H, W, Cc, heads, n_kv, dtype=torch.float32, win=7, with_pad_mask=False,
              eps=1e-6):
    return AttnDesc(B, H, W, Cc, heads, win, n_kv, _DT[dtype], int(with_pad_mask), eps)


def pack_attn(Cc, heads, win, ln_q, ln_kv, wq, bq, wk, bk, wv, bv, wo, bo, rpb_table):
    """-> fp32 CPU blob.  ln_q / ln_kv are (weight, bias) pairs."""
    lib = _lib.load()
    d = AttnDesc(1, win, win, Cc, heads, win, 0, HRF_F32, 0, 1e-6)
    blob = torch.empty(lib.hrf_attn_blob_floats(C.byref(d)), dtype=torch.float32)
    ts = [_host(t) if t is not None else None
          for t in (ln_q[0], ln_q[1], ln_kv[0], ln_kv[1], wq, bq, wk, bk, wv, bv, wo, bo, rpb_table)]
    check(lib.hrf_attn_pack(C.byref(d), *[_fp(t) for t in ts], _fp(blob)))
    return blob


def pack_ffn(ln, conv1, bn1, convd, bn2, conv2, bn3, bn_eps=1e-5):
    lib = _lib.load()
    Cc, hidden = conv1.weight.shape[1], conv1.weight.shape[0]
    d = FfnDesc(1, 8, 8, Cc, hidden, HRF_F32, 1e-6)
    blob = torch.empty(lib.hrf_ffn_blob_floats(C.byref(d)), dtype=torch.float32)
    keep = []
    h = lambda t: (keep.append(_host(t)) or keep[-1]) if t is not None else None
    args = [h(ln.weight), h(ln.bias), h(conv1.weight), h(conv1.bias), _bn4(bn1, keep),
            h(convd.weight), h(convd.bias), _bn4(bn2, keep), h(conv2.weight), h(conv2.bias),
            _bn4(bn3, keep)]
    conv = [a if isinstance(a, (type(None), C.Array)) else _fp(a) for a in args]
    check(lib.hrf_ffn_pack(C.byref(d), *conv, C.c_float(bn_eps), _fp(blob)))
    return blob


def pack_pw(conv, bn, bn_eps=1e-5):
    lib = _lib.load()
    cout, cin = conv.weight.shape[:2]
    d = PwDesc(1, 1, 1, cin, cout, HRF_F32, 0)
    blob = torch.empty(lib.hrf_pw_blob_floats(C.byref(d)), dtype=torch.float32)
    keep = []
    w = _host(conv.weight)
    b = _host(conv.bias) if conv.bias is not None else None
    check(lib.hrf_pw_pack(C.byref(d), _fp(w), _fp(b), _bn4(bn, keep), C.c_float(bn_eps), _fp(blob)))
    return blob


def pack_conv3x3(conv, bn, bn_eps=1e-5):
    """Dense 3x3 conv (pad 1) + BN folded -> blob of hrf_conv3x3_fwd."""
    lib = _lib.load()
    cout, cin = conv.weight.shape[:2]
    assert conv.weight.shape[2:] == (3, 3) and conv.groups == 1 and conv.padding == (1, 1)
    d = ConvDesc(1, 1, 1, cin, cout, conv.stride[0], HRF_F32, 0)
    blob = torch.empty(lib.hrf_conv3x3_blob_floats(C.byref(d)), dtype=torch.float32)
    keep = []
    w = _host(conv.weight)
    b = _host(conv.bias) if conv.bias is not None else None
    check(lib.hrf_conv3x3_pack(C.byref(d), _fp(w), _fp(b), _bn4(bn, keep), C.c_float(bn_eps), _fp(blob)))
    return blob


def convgemm_supported(cin, cout, ksize, stride):
    """True when hrf_convgemm_fwd covers a (cin -> cout, ksize, stride) convolution."""
    lib = _lib.load()
    d = ConvGemmDesc(1, 8, 16, cin, cout, ksize, stride, 0)
    return lib.hrf_convgemm_supported(C.byref(d)) == 0


def pack_convgemm(conv, bn, bn_eps=1e-5, extra_bias=None):
    """Dense 1x1 / 3x3 conv + BN folded (+ an extra per-channel bias) -> blob of hrf_convgemm_fwd."""
    lib = _lib.load()
    cout, cin, k, _ = conv.weight.shape
    assert conv.groups == 1 and conv.padding == (k // 2, k // 2) and conv.dilation == (1, 1)
    d = ConvGemmDesc(1, 8, 16, cin, cout, k, conv.stride[0], 0)
    n = lib.hrf_convgemm_blob_floats(C.byref(d))
    if n == 0:
        raise _lib.HrfError(f'convgemm: {k}x{k} conv {cin} -> {cout} stride {conv.stride[0]} is not covered')
    blob = torch.empty(n, dtype=torch.float32)
    keep = []
    w = _host(conv.weight)
    b = _host(conv.bias) if conv.bias is not None else None
    eb = _host(extra_bias) if extra_bias is not None else None
    check(lib.hrf_convgemm_pack(C.byref(d), _fp(w), _fp(b), _bn4(bn, keep), C.c_float(bn_eps), _fp(eb), _fp(blob)))
    return blob


def conv_gemm(x, blob, cout, ksize, stride=1, relu=False, residual=None):
    """bf16 tokens [B,H,W,Cin] -> [B,ceil(H/s),ceil(W/s),cout]: conv + folded BN (+ residual) (+ ReLU)
    on the warp-specialised TMA / tcgen05 implicit-GEMM kernel."""
    lib = _lib.load()
    _check_act(x)
    if x.dtype != torch.bfloat16:
        raise TypeError('conv_gemm runs in the bf16 mode only')
    B, H, W, Cin = x.shape
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    out = x.new_empty(B, Ho, Wo, cout)
    if residual is not None:
        if tuple(residual.shape) != tuple(out.shape) or residual.dtype != x.dtype or not residual.is_contiguous():
            raise ValueError('conv_gemm: residual must be a contiguous tensor of the output shape')
    d = ConvGemmDesc(B, H, W, Cin, cout, ksize, stride, int(relu))
    no = B * Ho * Wo
    with _timed('convgemm', C=Cin, launches=1,
                bytes=float((B * H * W * Cin + no * cout * (2 if residual is not None else 1)) * 2),
                flops=float(2 * no * ksize * ksize * Cin * cout)):
        check(lib.hrf_convgemm_fwd(C.byref(d), x.data_ptr(), residual.data_ptr() if residual is not None else None,
                                   blob.data_ptr(), out.data_ptr(), _stream()))
    return out


def conv_gemm_grouped(xs, blobs, cout, ksize, stride=1, relu=False, residuals=None):
    """`conv_gemm` of ONE layer shape on several tensors (the camera and modality streams), each
    with its own weights, in one launch (hrf_convgemm_grouped_fwd)."""
    lib = _lib.load()
    n = len(xs)
    for x in xs:
        _check_act(x)
        if x.dtype != torch.bfloat16 or tuple(x.shape) != tuple(xs[0].shape):
            raise ValueError('conv_gemm_grouped: bf16 tensors of one shape')
    B, H, W, Cin = xs[0].shape
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    outs = [x.new_empty(B, Ho, Wo, cout) for x in xs]
    if residuals is not None:
        for r in residuals:
            if tuple(r.shape) != tuple(outs[0].shape) or r.dtype != torch.bfloat16 or not r.is_contiguous():
                raise ValueError('conv_gemm_grouped: residuals must be contiguous tensors of the output shape')
    mask = sum(1 << q for q, r in enumerate(relu) if r) if isinstance(relu, (list, tuple)) else \
        ((1 << n) - 1 if relu else 0)
    d = ConvGemmDesc(B, H, W, Cin, cout, ksize, stride, mask)
    VP = C.c_void_p * n
    no = B * Ho * Wo
    with _timed('convgemm', C=Cin, launches=1,
                bytes=float(n * (B * H * W * Cin + no * cout * (2 if residuals is not None else 1)) * 2),
                flops=float(n * 2 * no * ksize * ksize * Cin * cout)):
        check(lib.hrf_convgemm_grouped_fwd(
            C.byref(d), n, VP(*[x.data_ptr() for x in xs]),
            VP(*[r.data_ptr() for r in residuals]) if residuals is not None else None,
            VP(*[b.data_ptr() for b in blobs]), VP(*[o.data_ptr() for o in outs]), _stream()))
    return outs


def conv_gemm_cat_grouped(xs, xs2, blobs, cout, relu=False):
    """1x1 conv of the channel concatenation [x | x2] (never materialised) on several tensors in
    one launch (hrf_convgemm_grouped_cat_fwd): Bottleneck conv3 + downsample as ONE GEMM."""
    lib = _lib.load()
    n = len(xs)
    for x, x2 in zip(xs, xs2):
        _check_act(x)
        _check_act(x2)
        if (x.dtype != torch.bfloat16 or x2.dtype != torch.bfloat16 or tuple(x.shape) != tuple(xs[0].shape) or
                tuple(x2.shape) != tuple(xs2[0].shape) or x.shape[:3] != x2.shape[:3]):
            raise ValueError('conv_gemm_cat_grouped: bf16 tensors, one shape per source, same B / H / W')
    B, H, W, c1 = xs[0].shape
    cin = c1 + xs2[0].shape[3]
    outs = [x.new_empty(B, H, W, cout) for x in xs]
    mask = sum(1 << q for q, r in enumerate(relu) if r) if isinstance(relu, (list, tuple)) else \
        ((1 << n) - 1 if relu else 0)
    d = ConvGemmDesc(B, H, W, cin, cout, 1, 1, mask)
    VP = C.c_void_p * n
    no = B * H * W
    with _timed('convgemm', C=cin, launches=1, bytes=float(n * no * (cin + cout) * 2),
                flops=float(n * 2 * no * cin * cout)):
        check(lib.hrf_convgemm_grouped_cat_fwd(
            C.byref(d), n, VP(*[x.data_ptr() for x in xs]), VP(*[x.data_ptr() for x in xs2]), c1,
            VP(*[b.data_ptr() for b in blobs]), VP(*[o.data_ptr() for o in outs]), _stream()))
    return outs


def pack_dwpw(conv_dw, bn_dw, conv_pw, bn_pw, bn_eps=1e-5):
    lib = _lib.load()
    cout, cin = conv_pw.weight.shape[:2]
    d = DwPwDesc(1, 2, 2, cin, cout, HRF_F32, 0)
    blob = torch.empty(lib.hrf_dwpw_blob_floats(C.byref(d)), dtype=torch.float32)
    keep = []
    wd, wp = _host(conv_dw.weight), _host(conv_pw.weight)
    check(lib.hrf_dwpw_pack(C.byref(d), _fp(wd), _bn4(bn_dw, keep), _fp(wp), _bn4(bn_pw, keep),
                            C.c_float(bn_eps), _fp(blob)))
    return blob


def pack_stem(conv, bn, bn_eps=1e-5):
    lib = _lib.load()
    cout, cin = conv.weight.shape[:2]
    d = StemDesc(1, cin, 2, 2, cout, 1)
    blob = torch.empty(lib.hrf_stem_blob_floats(C.byref(d)), dtype=torch.float32)
    keep = []
    w = _host(conv.weight)
    check(lib.hrf_stem_pack(C.byref(d), _fp(w), _bn4(bn, keep), C.c_float(bn_eps), _fp(blob)))
    return blob


# ----------------------------------------------------------------------------
# forward ops (device)
# ----------------------------------------------------------------------------
def stem_conv(x_nchw, blob, cout, relu=True):
    """fp32 NCHW image -> bf16 channels-last (B, H/2, W/2, cout): conv3x3 s2 + BN + ReLU"""
    lib = _lib.load()
    assert x_nchw.is_cuda and x_nchw.dtype == torch.float32 and x_nchw.is_contiguous()
    B, cin, H, W = x_nchw.shape
    out = torch.empty(B, (H + 1) // 2, (W + 1) // 2, cout, dtype=torch.bfloat16, device=x_nchw.device)
    d = StemDesc(B, cin, H, W, cout, int(relu))
    with _timed('stem_conv', C=cin, launches=1,
                bytes=float(x_nchw.numel() * 4 + out.numel() * 2),
                flops=float(out.numel() * 2 * 9 * cin)):
        check(lib.hrf_stem_conv_fwd(C.byref(d), x_nchw.data_ptr(), blob.data_ptr(), out.data_ptr(),
                                    _stream()))
    return out



def window_attention(x, kv, blobs, heads, win=7, with_pad_mask=False, eps=1e-6, out=None):
    """LSA when `kv` is empty/None (out = x + Attn(LN x)), else MWCA over the
    modalities in `kv` (out = x + sum_k[kv_k + Attn_k(LN1_k x, LN2_k kv_k)])."""
    lib = _lib.load()
    _check_act(x)
    kv = list(kv or [])
    for t in kv:
        _check_act(t, 'kv')
        assert t.shape == x.shape and t.dtype == x.dtype
    assert len(blobs) == max(1, len(kv))
    B, H, W, Cc = x.shape
    out = torch.empty_like(x) if out is None else out
    d = attn_desc(B, H, W, Cc, heads, len(kv), x.dtype, win, with_pad_mask, eps)
    # algorithmic cost: per token and modality pass,
    # FLOPs 8C^2 + 4*S*C (S = win^2 keys), bytes 2C (LSA) / 3C (MWCA) elements
    n, S, s = B * H * W, win * win, x.element_size()
    passes = max(1, len(kv))
    with _timed('mwca' if kv else 'lsa', C=Cc, launches=passes,
                bytes=float(n * passes * (3 if kv else 2) * Cc * s),
                flops=float(n * passes * (8 * Cc * Cc + 4 * S * Cc))):
        ws_bytes = lib.hrf_attn_workspace_bytes(C.byref(d))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if ws_bytes else None
        check(lib.hrf_window_attn_fwd(C.byref(d), x.data_ptr(), _ptr_array(kv), _ptr_array(blobs),
                                      out.data_ptr(), ws.data_ptr() if ws_bytes else None,
                                      ws_bytes, _stream()))
    return out


def window_attention_grouped(xs, blobs, heads, win=7, with_pad_mask=False, eps=1e-6):
    """LSA (out = x + Attn(LN x)) of several tensors of one shape, each with its own blob -- the
    camera's branch 0 and the modality streams -- through hrf_window_attn_grouped_fwd (one launch
    where the kernel covers it)."""
    lib = _lib.load()
    for x in xs:
        _check_act(x)
        assert x.shape == xs[0].shape and x.dtype == xs[0].dtype
    assert len(blobs) == len(xs)
    B, H, W, Cc = xs[0].shape
    outs = [torch.empty_like(x) for x in xs]
    d = attn_desc(B, H, W, Cc, heads, 0, xs[0].dtype, win, with_pad_mask, eps)
    n, S, s, q = B * H * W, win * win, xs[0].element_size(), len(xs)
    with _timed('lsa', C=Cc, launches=1, bytes=float(q * n * 2 * Cc * s), flops=float(q * n * (8 * Cc * Cc + 4 * S * Cc)),
                problems=q):
        ws_bytes = lib.hrf_attn_workspace_bytes(C.byref(d))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xs[0].device) if ws_bytes else None
        check(lib.hrf_window_attn_grouped_fwd(C.byref(d), q, _ptr_array(xs), _ptr_array(blobs), _ptr_array(outs),
                                              ws.data_ptr() if ws_bytes else None, ws_bytes, _stream()))
    return outs


def mixffn(x, blob, hidden, eps=1e-6, out=None):
    lib = _lib.load()
    _check_act(x)
    B, H, W, Cc = x.shape
    out = torch.empty_like(x) if out is None else out
    d = FfnDesc(B, H, W, Cc, hidden, _dtype_code(x), eps)
    n = B * H * W
    with _timed('mixffn', C=Cc, launches=1, bytes=float(n * 2 * Cc * x.element_size()),
                flops=float(n * (4 * Cc * hidden + 18 * hidden))):
        ws_bytes = lib.hrf_ffn_workspace_bytes(C.byref(d))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if ws_bytes else None
        check(lib.hrf_mixffn_fwd(C.byref(d), x.data_ptr(), blob.data_ptr(), out.data_ptr(),
                                 ws.data_ptr() if ws_bytes else None, ws_bytes, _stream()))
    return out


def mixffn_grouped(xs, blobs, hidden, eps=1e-6):
    """MixFFN block of several tensors of one shape, each with its own blob, through
    hrf_mixffn_grouped_fwd (one launch where the kernel covers it)."""
    lib = _lib.load()
    for x in xs:
        _check_act(x)
        assert x.shape == xs[0].shape and x.dtype == xs[0].dtype
    B, H, W, Cc = xs[0].shape
    outs = [torch.empty_like(x) for x in xs]
    d = FfnDesc(B, H, W, Cc, hidden, _dtype_code(xs[0]), eps)
    n, q = B * H * W, len(xs)
    with _timed('mixffn', C=Cc, launches=1, bytes=float(q * n * 2 * Cc * xs[0].element_size()),
                flops=float(q * n * (4 * Cc * hidden + 18 * hidden)), problems=q):
        ws_bytes = lib.hrf_ffn_workspace_bytes(C.byref(d))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xs[0].device) if ws_bytes else None
        check(lib.hrf_mixffn_grouped_fwd(C.byref(d), q, _ptr_array(xs), _ptr_array(blobs), _ptr_array(outs),
                                         ws.data_ptr() if ws_bytes else None, ws_bytes, _stream()))
    return outs


def pointwise(x, blob, cout, relu=False):
    lib = _lib.load()
    _check_act(x)
    B, H, W, Cin = x.shape
    out = x.new_empty(B, H, W, cout)
    d = PwDesc(B, H, W, Cin, cout, _dtype_code(x), int(relu))
    n = B * H * W
    with _timed('pw', C=Cin, launches=1, bytes=float(n * (Cin + cout) * x.element_size()),
                flops=float(2 * n * Cin * cout)):
        check(lib.hrf_pw_fwd(C.byref(d), x.data_ptr(), blob.data_ptr(), out.data_ptr(), _stream()))
    return out


def conv3x3(x, blob, cout, stride=1, relu=False):
    """tokens [B,H,W,Cin] -> tokens [B,ceil(H/s),ceil(W/s),cout] (3x3, pad 1, BN folded)."""
    lib = _lib.load()
    _check_act(x)
    B, H, W, Cin = x.shape
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    out = x.new_empty(B, Ho, Wo, cout)
    d = ConvDesc(B, H, W, Cin, cout, stride, _dtype_code(x), int(relu))
    no = B * Ho * Wo
    with _timed('conv3x3', C=Cin, launches=1,
                bytes=float((B * H * W * Cin + no * cout) * x.element_size()),
                flops=float(2 * no * 9 * Cin * cout)):
        check(lib.hrf_conv3x3_fwd(C.byref(d), x.data_ptr(), blob.data_ptr(), out.data_ptr(), _stream()))
    return out


def dw_down(x, blob, cout, relu=False):
    lib = _lib.load()
    _check_act(x)
    B, H, W, Cin = x.shape
    out = x.new_empty(B, (H + 1) // 2, (W + 1) // 2, cout)
    d = DwPwDesc(B, H, W, Cin, cout, _dtype_code(x), int(relu))
    no = out.shape[0] * out.shape[1] * out.shape[2]
    with _timed('dwpw', C=Cin, launches=1,
                bytes=float((B * H * W * Cin + no * cout) * x.element_size()),
                flops=float(no * (18 * Cin + 2 * Cin * cout))):
        check(lib.hrf_dwpw_fwd(C.byref(d), x.data_ptr(), blob.data_ptr(), out.data_ptr(),
                               _stream()))
    return out


def fuse_sum(x, ups=(), sames=(), relu=True, nchw_out=False):
    """ReLU(x + sum bilinear(ups) + sum sames); with nchw_out also returns the
    result as a contiguous fp32 (B,C,H,W) tensor."""
    lib = _lib.load()
    _check_act(x)
    B, H, W, Cc = x.shape
    ups, sames = list(ups), list(sames)
    d = FuseDesc(B, H, W, Cc, _dtype_code(x), len(ups))
    for j, u in enumerate(ups):
        _check_act(u, 'up')
        assert u.shape[0] == B and u.shape[3] == Cc and u.dtype == x.dtype
        d.up_H[j], d.up_W[j] = u.shape[1], u.shape[2]
    for s in sames:
        _check_act(s, 'same')
        assert s.shape == x.shape and s.dtype == x.dtype
    d.n_same, d.relu = len(sames), int(relu)
    out = torch.empty_like(x)
    nchw = torch.empty(B, Cc, H, W, dtype=torch.float32, device=x.device) if nchw_out else None
    s_ = x.element_size()
    nbytes = x.numel() * s_ * (2 + len(sames)) + sum(u.numel() for u in ups) * s_ + \
        (x.numel() * 4 if nchw_out else 0)
    with _timed('fuse_sum', C=Cc, launches=1, bytes=float(nbytes),
                flops=float(x.numel() * (1 + len(sames) + 8 * len(ups)))):
        check(lib.hrf_fuse_sum_fwd(C.byref(d), x.data_ptr(), _ptr_array(ups), _ptr_array(sames),
                                   out.data_ptr(), nchw.data_ptr() if nchw_out else None,
                                   _stream()))
    return (out, nchw) if nchw_out else out


def bias_act_(y, bias, residual=None, relu=True):
    """in place: y = act(y + bias[c] (+ residual)) on a channels-last (B,H,W,C) tensor"""
    lib = _lib.load()
    _check_act(y, 'y')
    B, H, W, Cc = y.shape
    assert bias.dtype == torch.float32 and bias.numel() == Cc
    if residual is not None:
        _check_act(residual, 'residual')
        assert residual.shape == y.shape and residual.dtype == y.dtype
    with _timed('bias_act', C=Cc, launches=1,
                bytes=float(y.numel() * y.element_size() * (3 if residual is not None else 2)),
                flops=float(y.numel() * 2)):
        check(lib.hrf_bias_act_fwd(B * H * W, Cc, _dtype_code(y), int(relu), y.data_ptr(),
                                   bias.data_ptr(), residual.data_ptr() if residual is not None
                                   else None, _stream()))
    return y


def input_prologue(src, mean, std, to_rgb=False, size_divisor=32, pad_val=0.0, out=None):
    """(B, H, W, C) uint8 / fp32 sensor frames on the device -> normalised, padded fp32 NCHW
    (B, C, ceil(H/div)*div, ceil(W/div)*div): `Normalize` + `Pad(size_divisor)` +
    `DefaultFormatBundle` + collate of the reference's data pipeline in one kernel
    (transforms.py:652-667,719-744, formating.py:211-227)."""
    lib = _lib.load()
    if not (src.is_cuda and src.dim() == 4 and src.is_contiguous()):
        raise ValueError('src must be a contiguous CUDA (B,H,W,C) tensor')
    if src.dtype not in (torch.uint8, torch.float32):
        raise TypeError(f'unsupported sensor frame dtype {src.dtype}')
    B, H, W, Cc = src.shape
    mean = [float(m) for m in mean]
    std = [float(v) for v in std]
    if len(mean) != Cc or len(std) != Cc:
        raise ValueError(f'mean / std must have {Cc} entries')
    div = int(size_divisor) if size_divisor else 1
    Hp, Wp = -(-H // div) * div, -(-W // div) * div
    if out is None:
        out = torch.empty(B, Cc, Hp, Wp, dtype=torch.float32, device=src.device)
    elif tuple(out.shape) != (B, Cc, Hp, Wp) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError('out must be a contiguous fp32 (B,C,Hp,Wp) tensor')
    d = InputDesc(B, H, W, Cc, Hp, Wp, HRF_U8 if src.dtype == torch.uint8 else HRF_F32, int(bool(to_rgb)),
                  float(pad_val))
    m = (C.c_float * Cc)(*mean)
    s = (C.c_float * Cc)(*std)
    with _timed('input_prologue', C=Cc, launches=1,
                bytes=float(src.numel() * src.element_size() + out.numel() * 4), flops=float(2 * src.numel())):
        check(lib.hrf_input_prologue_fwd(C.byref(d), src.data_ptr(), m, s, out.data_ptr(), _stream()))
    return out


def pool_tokens(x, k, mode='AVG'):
    """bf16 tokens [B,H,W,C] -> [B,H/k,W/k,C]: k x k pooling with stride k (HRFPN pyramid)."""
    lib = _lib.load()
    _check_act(x)
    if x.dtype != torch.bfloat16:
        raise TypeError('pool_tokens runs in the bf16 mode only')
    B, H, W, Cc = x.shape
    out = x.new_empty(B, H // k, W // k, Cc)
    with _timed('pool', C=Cc, launches=1, bytes=float((x.numel() + out.numel()) * 2), flops=float(x.numel())):
        check(lib.hrf_pool_fwd(B, H, W, Cc, k, int(mode == 'MAX'), x.data_ptr(), out.data_ptr(), _stream()))
    return out


def nchw_to_nhwc(x, dtype=None):
    lib = _lib.load()
    assert x.is_cuda and x.dim() == 4 and x.is_contiguous()
    B, Cc, H, W = x.shape
    out = torch.empty(B, H, W, Cc, dtype=dtype or x.dtype, device=x.device)
    check(lib.hrf_nchw_to_nhwc(B, Cc, H, W, _dtype_code(x), x.data_ptr(), _dtype_code(out),
                               out.data_ptr(), _stream()))
    return out


def nhwc_to_nchw(x, dtype=None):
    lib = _lib.load()
    _check_act(x)
    B, H, W, Cc = x.shape
    out = torch.empty(B, Cc, H, W, dtype=dtype or x.dtype, device=x.device)
    check(lib.hrf_nhwc_to_nchw(B, Cc, H, W, _dtype_code(x), x.data_ptr(), _dtype_code(out),
                               out.data_ptr(), _stream()))
    return out


def launch_count():
    return int(_lib.load().hrf_launch_count())


def set_pdl(enable):
    """Programmatic dependent launch on/off (hrf_set_pdl); returns the previous setting."""
    return bool(_lib.load().hrf_set_pdl(1 if enable else 0))


# ----------------------------------------------------------------------------
# train-mode BatchNorm statistics / affine passes (NCHW planes)
# ----------------------------------------------------------------------------
def _bn_desc(x):
    if not (x.is_cuda and x.dim() >= 2 and x.is_contiguous()):
        raise ValueError('x must be a contiguous CUDA (B, C, ...) tensor')
    B, Cc = x.shape[0], x.shape[1]
    hw = x.numel() // (B * Cc)
    return _lib.BnDesc(B, Cc, hw, _dtype_code(x))


_BN_WS = {}


def _bn_workspace(lib, d, device):
    """Partials buffer, cached per (device, stream): the calls on one stream are ordered."""
    need = lib.hrf_bn_workspace_bytes(C.byref(d))
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    ws = _BN_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = _BN_WS[key] = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=device)
    return ws


def _f32c(t, name):
    if t is None:
        return None
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise ValueError(f'{name} must be a contiguous fp32 CUDA tensor')
    return t.data_ptr()


def _like_x(dy, x):
    if dy.shape != x.shape or dy.dtype != x.dtype or not dy.is_contiguous():
        raise ValueError('dy must match x (shape, dtype, contiguous)')


def attn_core_train_supported(N, hd):
    return N <= 64 and hd <= 64


def attn_core_train_fwd(q, k, v, heads, scale, table=None, rpi=None):
    """softmax(scale q k^T + table[rpi]) v per (window, head); q / k / v: contiguous fp32 CUDA
    [nWin, N, C] -> (o [nWin, N, C], P [nWin, heads, N, N])."""
    lib = _lib.load()
    nWin, N, Cc = q.shape
    o = torch.empty_like(q)
    P = torch.empty(nWin, heads, N, N, dtype=torch.float32, device=q.device)
    with _timed('attn_core_train_fwd', C=Cc, bytes=4 * q.numel() * 4, flops=4.0 * nWin * N * N * Cc):
        check(lib.hrf_attn_core_train_fwd(nWin, N, Cc, heads, scale, q.data_ptr(), k.data_ptr(), v.data_ptr(),
                                          table.data_ptr() if table is not None else None,
                                          rpi.data_ptr() if table is not None else None,
                                          o.data_ptr(), P.data_ptr(), _stream()))
    return o, P


def attn_core_train_bwd(q, k, v, P, dout, heads, scale, rpi=None, T=0):
    """-> (dq, dk, dv, dtable [T, heads] | None)"""
    lib = _lib.load()
    nWin, N, Cc = q.shape
    dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
    dtable = ws = None
    n_ws = 0
    if rpi is not None:
        dtable = torch.empty(T, heads, dtype=torch.float32, device=q.device)
        n_ws = lib.hrf_attn_core_train_ws_floats(nWin, N, heads)
        ws = torch.empty(n_ws, dtype=torch.float32, device=q.device)
    with _timed('attn_core_train_bwd', C=Cc, bytes=8 * q.numel() * 4, flops=10.0 * nWin * N * N * Cc):
        check(lib.hrf_attn_core_train_bwd(nWin, N, Cc, heads, scale, q.data_ptr(), k.data_ptr(), v.data_ptr(),
                                          P.data_ptr(), dout.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                          rpi.data_ptr() if rpi is not None else None, T,
                                          dtable.data_ptr() if dtable is not None else None,
                                          ws.data_ptr() if ws is not None else None, n_ws, _stream()))
    return dq, dk, dv, dtable


def dwconv_train_fwd(x, weight, bias, stride):
    """depthwise 3x3 (pad 1) of a contiguous fp32 CUDA NCHW tensor; weight (C, 1, 3, 3)"""
    lib = _lib.load()
    B, Cc, H, W = x.shape
    y = torch.empty(B, Cc, (H - 1) // stride + 1, (W - 1) // stride + 1, dtype=torch.float32, device=x.device)
    with _timed('dwconv_train_fwd', C=Cc, bytes=(x.numel() + y.numel()) * 4, flops=18.0 * y.numel()):
        check(lib.hrf_dwconv_train_fwd(B, Cc, H, W, stride, x.data_ptr(), weight.data_ptr(),
                                       bias.data_ptr() if bias is not None else None, y.data_ptr(), _stream()))
    return y


def dwconv_train_bwd(x, g, weight, stride, want_dx=True, want_dw=True, want_dbias=False):
    """-> (dx | None, dweight (C, 1, 3, 3) | None, dbias | None)"""
    lib = _lib.load()
    B, Cc, H, W = x.shape
    dx = dw = db = None
    if want_dx:
        dx = torch.empty_like(x)
        with _timed('dwconv_train_dgrad', C=Cc, bytes=(x.numel() + g.numel()) * 4, flops=18.0 * g.numel()):
            check(lib.hrf_dwconv_train_dgrad(B, Cc, H, W, stride, g.data_ptr(), weight.data_ptr(), dx.data_ptr(),
                                             _stream()))
    if want_dw or want_dbias:
        dw = torch.empty_like(weight)
        db = torch.empty(Cc, dtype=torch.float32, device=x.device) if want_dbias else None
        n_ws = lib.hrf_dwconv_train_ws_floats(B, Cc, H, W, stride)
        ws = torch.empty(n_ws, dtype=torch.float32, device=x.device)
        with _timed('dwconv_train_wgrad', C=Cc, bytes=(x.numel() + g.numel()) * 4, flops=18.0 * g.numel()):
            check(lib.hrf_dwconv_train_wgrad(B, Cc, H, W, stride, x.data_ptr(), g.data_ptr(), dw.data_ptr(),
                                             db.data_ptr() if db is not None else None, ws.data_ptr(), n_ws,
                                             _stream()))
    return dx, dw, db


def ln_fwd(x, weight, bias, eps):
    """Train-mode LayerNorm over the last axis of a contiguous fp32 CUDA tensor -> (y, mean, rstd)."""
    lib = _lib.load()
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    y = torch.empty_like(x)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty_like(mean)
    with _timed('ln_fwd', C=Cc, bytes=2 * x.numel() * 4, flops=0.0):
        check(lib.hrf_ln_fwd(rows, Cc, eps, x.data_ptr(), weight.data_ptr(), bias.data_ptr(), y.data_ptr(),
                             mean.data_ptr(), rstd.data_ptr(), _stream()))
    return y, mean, rstd


def ln_bwd(x, dy, mean, rstd, weight, want_dx=True):
    """-> (dx | None, dweight, dbias) of `ln_fwd`."""
    lib = _lib.load()
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    dx = torch.empty_like(x) if want_dx else None
    dw = torch.empty(Cc, dtype=torch.float32, device=x.device)
    db = torch.empty_like(dw)
    n_ws = lib.hrf_ln_bwd_workspace_floats(rows, Cc)
    ws = torch.empty(n_ws, dtype=torch.float32, device=x.device)
    with _timed('ln_bwd', C=Cc, bytes=3 * x.numel() * 4, flops=0.0):
        check(lib.hrf_ln_bwd(rows, Cc, x.data_ptr(), dy.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                             weight.data_ptr(), dx.data_ptr() if want_dx else None, dw.data_ptr(), db.data_ptr(),
                             ws.data_ptr(), n_ws, _stream()))
    return dx, dw, db


def bn_stats(x):
    """x (B, C, *) contiguous -> fp64 [2C + 1]: per-channel sum(x) | sum(x^2) | element count.
    The SyncBN message: additive across ranks."""
    lib = _lib.load()
    d = _bn_desc(x)
    ws = _bn_workspace(lib, d, x.device)
    stats = torch.empty(2 * d.C + 1, dtype=torch.float64, device=x.device)
    with _timed('bn_stats', C=d.C, bytes=x.numel() * x.element_size(), flops=0.0):
        check(lib.hrf_bn_stats(C.byref(d), x.data_ptr(), stats.data_ptr(), ws.data_ptr(),
                               ws.numel(), _stream()))
    return stats


BN_ACT_NONE, BN_ACT_RELU, BN_ACT_GELU = 0, 1, 2


def bn_normalize(x, stats, weight, bias, eps, momentum=0.0, running_mean=None, running_var=None,
                 relu=False, act=BN_ACT_NONE):
    """y = act(BN(x)) from the (all-reduced) `bn_stats` message; -> (y, mean, invstd); updates the
    running statistics in place when given.  act: BN_ACT_NONE | BN_ACT_RELU | BN_ACT_GELU
    (`relu=True` is BN_ACT_RELU)."""
    act = BN_ACT_RELU if relu and not act else int(act)
    lib = _lib.load()
    d = _bn_desc(x)
    y = torch.empty_like(x)
    mean = torch.empty(d.C, dtype=torch.float32, device=x.device)
    invstd = torch.empty_like(mean)
    with _timed('bn_affine', C=d.C, bytes=2 * x.numel() * x.element_size(), flops=0.0):
        check(lib.hrf_bn_normalize(C.byref(d), x.data_ptr(), stats.data_ptr(), _f32c(weight, 'weight'),
                                   _f32c(bias, 'bias'), float(eps), float(momentum),
                                   _f32c(running_mean, 'running_mean'),
                                   _f32c(running_var, 'running_var'), mean.data_ptr(),
                                   invstd.data_ptr(), act, y.data_ptr(), _stream()))
    return y, mean, invstd


def bn_bwd_stats(x, dy, mean, invstd, want_param_grads=False, weight=None, bias=None, act=BN_ACT_NONE):
    """-> fp64 [2C]: per-channel sum(dy) | sum(dy * (x - mean) * invstd), and with
    `want_param_grads` the fp32 (dweight, dbias) copies of the same (rank-local) numbers.
    With `act`, dy is the gradient of act(BN(x)) and is multiplied by act'(z) on the fly
    (z recomputed from x, the statistics and weight / bias)."""
    lib = _lib.load()
    d = _bn_desc(x)
    _like_x(dy, x)
    ws = _bn_workspace(lib, d, x.device)
    sums = torch.empty(2 * d.C, dtype=torch.float64, device=x.device)
    grads = torch.empty(2, d.C, dtype=torch.float32, device=x.device) if want_param_grads else None
    with _timed('bn_bwd_stats', C=d.C, bytes=2 * x.numel() * x.element_size(), flops=0.0):
        check(lib.hrf_bn_bwd_stats(C.byref(d), x.data_ptr(), dy.data_ptr(), _f32c(mean, 'mean'),
                                   _f32c(invstd, 'invstd'), _f32c(weight, 'weight'),
                                   _f32c(bias, 'bias'), int(act), sums.data_ptr(),
                                   grads[0].data_ptr() if want_param_grads else None,
                                   grads[1].data_ptr() if want_param_grads else None,
                                   ws.data_ptr(), ws.numel(), _stream()))
    return (sums, grads[0], grads[1]) if want_param_grads else sums


def bn_bwd_dx(x, dy, sums, count, weight, mean, invstd, bias=None, act=BN_ACT_NONE):
    """dx of train-mode BN (followed by `act`) from the (all-reduced) `bn_bwd_stats` sums;
    `count` is the fp64 device scalar the forward message carried (stats[2C:])."""
    lib = _lib.load()
    d = _bn_desc(x)
    _like_x(dy, x)
    dx = torch.empty_like(x)
    with _timed('bn_bwd_affine', C=d.C, bytes=3 * x.numel() * x.element_size(), flops=0.0):
        check(lib.hrf_bn_bwd_dx(C.byref(d), x.data_ptr(), dy.data_ptr(), sums.data_ptr(),
                                count.data_ptr(), _f32c(weight, 'weight'), _f32c(bias, 'bias'),
                                int(act), _f32c(mean, 'mean'), _f32c(invstd, 'invstd'),
                                dx.data_ptr(), _stream()))
    return dx


def bn_affine(x, a, c0, dy=None, b=None, relu=False, out=None):
    """out = a[c]*x + c0[c]   or, with dy,  a[c]*dy + b[c]*x + c0[c]  (per-channel fp32 a, b, c0)."""
    lib = _lib.load()
    d = _bn_desc(x)
    out = torch.empty_like(x) if out is None else out
    a, c0 = a.float().contiguous(), c0.float().contiguous()
    if dy is not None:
        _like_x(dy, x)
        b = b.float().contiguous()
    with _timed('bn_affine' if dy is None else 'bn_bwd_affine', C=d.C,
                bytes=x.numel() * x.element_size() * (2 if dy is None else 3), flops=0.0):
        check(lib.hrf_bn_affine(C.byref(d), x.data_ptr(), dy.data_ptr() if dy is not None else None,
                                a.data_ptr(), b.data_ptr() if dy is not None else None,
                                c0.data_ptr(), int(relu), out.data_ptr(), _stream()))
    return out
