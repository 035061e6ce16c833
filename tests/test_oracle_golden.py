"""Pin the oracle to the committed outputs of the unmodified reference."""
import copy
import os

import numpy as np
import pytest
import torch

from helpers import BN, LN, rel_err
from hrfuser_b200 import HRFuserHRFormerBased, backbone_cfg
from hrfuser_b200.modules import HRFormerBlock, HRFormerModule, HRFuserFusionBlock
from hrfuser_b200.utils import randomize_parameters, synthetic_inputs
from oracle import hrfuser_oracle as O

G = os.path.join(os.path.dirname(__file__), 'golden')
OPS = np.load(os.path.join(G, 'ops.npz'))
E2E = np.load(os.path.join(G, 'e2e.npz'))
TOL = 2e-6          # fp32 reduction-order noise of the reference itself is ~3e-7 (App. F)


def _checksum(m):
    return float(sum(v.double().sum() for v in m.state_dict().values() if v.is_floating_point()))


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize('name', ['blk_c18', 'blk_c36', 'blk_c72'])
def test_hrformer_block(name):
    C, heads, H, W, B = (int(v) for v in OPS[name + '_cfg'])
    blk = HRFormerBlock(C, C, heads, 7, 4, 0., BN, LN)
    randomize_parameters(blk, 11)
    assert abs(_checksum(blk) - float(OPS[name + '_checksum'])) < 1e-6, 'parameter generator drifted'
    sd = {'blk.' + k: v for k, v in blk.state_dict().items()}
    x = _t(OPS[name + '_x'])
    t = x.flatten(2).transpose(1, 2)
    n = O.layer_norm(t, sd, 'blk.norm1')
    a = t + O.window_attention(n, n, sd, 'blk.attn', H, W, heads, cross=False)
    assert rel_err(a, _t(OPS[name + '_after_attn'])) < TOL
    assert rel_err(O.hrformer_block(x, sd, 'blk', heads), _t(OPS[name + '_y'])) < TOL


@pytest.mark.parametrize('name', ['fus_c36_m2', 'fus_c18_m3'])
def test_fusion_block(name):
    C, heads, H, W, M = (int(v) for v in OPS[name + '_cfg'])
    blk = HRFuserFusionBlock(C, C, heads, 7, 4, 0., BN, LN, num_fused_modalities=M, proj_drop_rate=0.1)
    randomize_parameters(blk, 12)
    assert abs(_checksum(blk) - float(OPS[name + '_checksum'])) < 1e-6
    sd = {'blk.' + k: v for k, v in blk.state_dict().items()}
    zs = [_t(OPS[f'{name}_z{k}']) for k in range(M)]
    y = O.fusion_block(_t(OPS[name + '_x']), zs, sd, 'blk', heads)
    assert rel_err(y, _t(OPS[name + '_y'])) < TOL


def test_exchange_module():
    mod = HRFormerModule(3, (1, 1, 1), [18, 36, 72], (1, 2, 4), (7, 7, 7), (4, 4, 4), True, BN, LN, [0.0])
    randomize_parameters(mod, 13)
    assert abs(_checksum(mod) - float(OPS['mod3_checksum'])) < 1e-6
    sd = {'s.0.' + k: v for k, v in mod.state_dict().items()}
    cfg = dict(num_modules=1, num_blocks=(1, 1, 1), num_heads=(1, 2, 4))
    ys = O.hrformer_stage([_t(OPS[f'mod3_x{i}']) for i in range(3)], sd, 's', cfg)
    for i in range(3):
        assert rel_err(ys[i], _t(OPS[f'mod3_y{i}'])) < TOL


def _net(v, d):
    cfg = backbone_cfg(v, d)
    c = copy.deepcopy(cfg)
    c.pop('type')
    net = HRFuserHRFormerBased(**c)
    randomize_parameters(net, 1)
    return cfg, net


@pytest.mark.parametrize('tag,v,d,mc', [('t_nus', 't', 'nus', (3, 3)), ('t_stf', 't', 'stf', (3, 2, 1)),
                                        ('b_nus', 'b', 'nus', (3, 3))])
def test_backbone_small_input_with_stage_taps(tag, v, d, mc):
    cfg, net = _net(v, d)
    assert abs(_checksum(net) - float(E2E[tag + '_checksum'])) < 1e-4
    H, W = (int(t) for t in E2E[tag + '_hw'])
    x, mods = synthetic_inputs(1, H, W, mc, seed=3)
    taps = {}
    ys = O.backbone_forward(net.state_dict(), cfg, x, mods, taps)
    for i, y in enumerate(ys):
        assert rel_err(y, _t(E2E[f'{tag}_out{i}'])) < TOL
    n = 0
    for stage in ('fusion_a', 'fusion_b', 'fusion_c', 'stage2', 'stage3', 'stage4'):
        for i, t in enumerate(taps[stage]):
            assert rel_err(t, _t(E2E[f'{tag}_{stage}.{i}'])) < TOL, (stage, i)
            n += 1
    assert n == 2 + 3 + 4 + 2 + 3 + 4


def test_backbone_full_size_statistics():
    """configs[0] (HRFuser-T nuScenes, batch 1, 384x640) against the reference's
    per-channel statistics and a fixed sample of 4096 values per output."""
    cfg, net = _net('t', 'nus')
    x, mods = synthetic_inputs(1, 384, 640, (3, 3), seed=0)
    with torch.no_grad():
        ys = O.backbone_forward(net.state_dict(), cfg, x, mods)
    for i, y in enumerate(ys):
        yd = y.double()
        assert rel_err(yd.mean((0, 2, 3)), _t(E2E[f'full_t_nus_out{i}_chan_mean'])) < 1e-5
        assert rel_err((yd ** 2).mean((0, 2, 3)), _t(E2E[f'full_t_nus_out{i}_chan_sqmean'])) < 1e-5
        idx = _t(E2E[f'full_t_nus_out{i}_idx'])
        assert rel_err(y.flatten()[idx], _t(E2E[f'full_t_nus_out{i}_val'])) < TOL


# ---- windows other than 7 (tests/golden/make_golden_win.py) ---------------------------------
OPS_WIN = np.load(os.path.join(G, 'ops_win.npz'))
WIN_CASES = ['w14_c36', 'w14_c18', 'w5_c36', 'w16_c18']


def _win_blocks(name):
    win, C, heads, H, W = (int(v) for v in OPS_WIN[name + '_cfg'])
    blk = HRFormerBlock(C, C, heads, win, 4, 0., BN, LN)
    randomize_parameters(blk, 21)
    fus = HRFuserFusionBlock(C, C, heads, win, 4, 0., BN, LN, num_fused_modalities=2, proj_drop_rate=0.1)
    randomize_parameters(fus, 22)
    assert abs(_checksum(blk) - float(OPS_WIN[name + '_checksum'])) < 1e-6
    assert abs(_checksum(fus) - float(OPS_WIN[name + '_fus_checksum'])) < 1e-6
    return win, C, heads, H, W, blk.eval(), fus.eval()


@pytest.mark.parametrize('name', WIN_CASES)
def test_window_sizes_oracle(name):
    """oracle window attention with Wh = Ww != 7 against the reference's blocks"""
    win, C, heads, H, W, blk, fus = _win_blocks(name)
    sd = {'blk.' + k: v for k, v in blk.state_dict().items()}
    x = _t(OPS_WIN[name + '_x'])
    t = x.flatten(2).transpose(1, 2)
    n = O.layer_norm(t, sd, 'blk.norm1')
    a = t + O.window_attention(n, n, sd, 'blk.attn', H, W, heads, cross=False, Wh=win, Ww=win)
    assert rel_err(a, _t(OPS_WIN[name + '_after_attn'])) < TOL
    y = a + O.cross_ffn(O.layer_norm(a, sd, 'blk.norm2'), sd, 'blk.ffn', H, W)
    assert rel_err(y.transpose(1, 2).reshape(x.shape), _t(OPS_WIN[name + '_y'])) < TOL
    fsd = {'blk.' + k: v for k, v in fus.state_dict().items()}
    acc = t.clone()
    for k in range(2):
        z = _t(OPS_WIN[f'{name}_z{k}']).flatten(2).transpose(1, 2)
        acc = acc + z + O.window_attention(O.layer_norm(t, fsd, f'blk.norm1.{k}'),
                                           O.layer_norm(z, fsd, f'blk.norm2.{k}'), fsd,
                                           f'blk.attn.{k}', H, W, heads, cross=True, Wh=win, Ww=win)
    acc = acc + O.cross_ffn(O.layer_norm(acc, fsd, 'blk.norm3'), fsd, 'blk.ffn', H, W)
    assert rel_err(acc.transpose(1, 2).reshape(x.shape), _t(OPS_WIN[name + '_fus_y'])) < TOL


@pytest.mark.parametrize('name', WIN_CASES)
def test_window_sizes_packers(built_lib, name):
    """the C packers + engine wiring for win != 7 (blob-level CPU emulation of the device ops)
    and the training-path modules against the same goldens"""
    import blob_emul
    from hrfuser_b200.engine import BackboneEngine
    win, C, heads, H, W, blk, fus = _win_blocks(name)
    e = BackboneEngine.__new__(BackboneEngine)
    e._host_blobs, e._blob_slots = [], []
    e.device, e.dtype = torch.device('cpu'), torch.float32
    p_lsa, p_fus = e._hrformer_block(blk), e._fusion_block(fus)
    e._upload()
    assert p_lsa['win'] == win
    x = _t(OPS_WIN[name + '_x'])
    xt = x.permute(0, 2, 3, 1).contiguous()
    a = blob_emul.window_attention(xt, None, [s.t for s in p_lsa['attn']], heads, win=win)
    assert rel_err(a.reshape(1, H * W, C), _t(OPS_WIN[name + '_after_attn'])) < 1e-5
    zs = [_t(OPS_WIN[f'{name}_z{k}']).permute(0, 2, 3, 1).contiguous() for k in range(2)]
    acc = blob_emul.window_attention(xt, zs, [s.t for s in p_fus['attn']], heads, win=win)
    f = p_fus['ffn']
    y = blob_emul.mixffn(acc, f['blob'].t, f['hidden'], f['eps'])
    assert rel_err(y.permute(0, 3, 1, 2), _t(OPS_WIN[name + '_fus_y'])) < 1e-5
    with torch.no_grad():
        assert rel_err(blk(x), _t(OPS_WIN[name + '_y'])) < TOL
        assert rel_err(fus(x, [_t(OPS_WIN[f'{name}_z{k}']) for k in range(2)]),
                       _t(OPS_WIN[name + '_fus_y'])) < TOL
