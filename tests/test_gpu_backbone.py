"""End-to-end GPU parity of the drop-in backbone against the oracle."""
import copy

import pytest
import torch

from helpers import assert_parity
from hrfuser_b200 import HRFuserHRFormerBased, backbone_cfg, tiny_cfg
from hrfuser_b200.utils import randomize_parameters, rel_err, synthetic_inputs
from oracle import hrfuser_oracle as O

pytestmark = pytest.mark.gpu


def _build(cfg, precision, seed=1):
    c = copy.deepcopy(cfg)
    c.pop('type')
    net = HRFuserHRFormerBased(**c, precision=precision)
    randomize_parameters(net, seed)
    net.eval()
    return net


def _run(cfg, mod_ch, B, H, W, precision, sparse=False, drop=None):
    net = _build(cfg, precision)
    x, mods = synthetic_inputs(B, H, W, mod_ch, seed=3, sparse=sparse)
    if drop is not None:
        mods[drop].zero_()
    with torch.no_grad():
        ref = O.backbone_forward(net.state_dict(), cfg, x, mods)
    net.cuda()
    with torch.no_grad():
        got = net(x.cuda(), [m.cuda() for m in mods])
    torch.cuda.synchronize()
    assert isinstance(got, list) and len(got) == 4
    return got, ref


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_tiny_topology(built_lib, precision):
    got, ref = _run(tiny_cfg(2), (3, 3), 2, 64, 96, precision)
    for i, (g, r) in enumerate(zip(got, ref)):
        assert g.dtype == torch.float32 and g.is_contiguous() and g.shape == r.shape
        if precision == 'fp32':
            assert_parity(g, r, 'fp32', f'out{i}')
        else:
            assert_parity(g, r, 'bf16', f'tiny out{i}')       # norm-wise <= 2e-2 and 99 % elementwise


def test_t_nus_full_fp32(built_lib):
    """configs[0]: HRFuser-T nuScenes, batch 1, 384x640, cam+lidar+radar."""
    got, ref = _run(backbone_cfg('t', 'nus'), (3, 3), 1, 384, 640, 'fp32')
    shapes = [(1, 18, 96, 160), (1, 36, 48, 80), (1, 72, 24, 40), (1, 144, 12, 20)]
    for i, (g, r) in enumerate(zip(got, ref)):
        assert tuple(g.shape) == shapes[i]
        assert_parity(g, r, 'fp32', f'T-nus out{i}')


def test_t_nus_full_bf16(built_lib):
    got, ref = _run(backbone_cfg('t', 'nus'), (3, 3), 1, 384, 640, 'bf16')
    for i, (g, r) in enumerate(zip(got, ref)):
        assert_parity(g, r, 'bf16', f'T-nus bf16 out{i}')   # north_star: 2e-2 in the bf16 mode


def test_t_stf_4mod_fp32_sparse_and_dropped(built_lib):
    """STF topology (3 extra modalities, 2- and 1-channel sensors), sparse
    projected-sensor-like inputs with one modality dropped to all zeros."""
    got, ref = _run(backbone_cfg('t', 'stf'), (3, 2, 1), 1, 96, 320, 'fp32', sparse=True, drop=1)
    for i, (g, r) in enumerate(zip(got, ref)):
        assert_parity(g, r, 'fp32', f'T-stf out{i}')


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_b_nus_all_widths(built_lib, precision):
    """HRFuser-B (C = 78 / 156 / 312 / 624, head_dim 39): fused SIMT kernels on the two
    high-resolution branches, the generic un-fused path on the wide ones."""
    got, ref = _run(backbone_cfg('b', 'nus'), (3, 3), 1, 64, 96, precision)
    for i, (g, r) in enumerate(zip(got, ref)):
        assert_parity(g, r, precision, f'B-nus out{i}')


def test_forward_spellings_and_errors(built_lib):
    net = _build(tiny_cfg(2), 'fp32').cuda()
    x, mods = synthetic_inputs(1, 32, 32, (3, 3), device='cuda')
    with torch.no_grad():
        a = net(x, mods)
        b = net(x, *mods)                       # forward(img, *extra_modalities)
    assert all(torch.equal(p, q) for p, q in zip(a, b))
    with pytest.raises(Exception, match='num_fused_modalities'):
        net(x, mods[:1])


def test_state_dict_roundtrip_changes_output(built_lib):
    """load_state_dict must invalidate the packed weights."""
    net = _build(tiny_cfg(2), 'fp32', seed=1).cuda()
    other = _build(tiny_cfg(2), 'fp32', seed=2)
    x, mods = synthetic_inputs(1, 32, 32, (3, 3), device='cuda')
    with torch.no_grad():
        a = net(x, mods)
        net.load_state_dict(other.state_dict())
        b = net(x, mods)
        ref = O.backbone_forward(other.state_dict(), tiny_cfg(2), x.cpu(), [m.cpu() for m in mods])
    assert not torch.equal(a[0], b[0])
    for i, (g, r) in enumerate(zip(b, ref)):
        assert_parity(g, r, 'fp32', f'reloaded out{i}')


def test_native_library_is_what_ran(built_lib):
    from hrfuser_b200 import ops
    net = _build(tiny_cfg(2), 'fp32').cuda()
    x, mods = synthetic_inputs(1, 32, 32, (3, 3), device='cuda')
    n0 = ops.launch_count()
    with torch.no_grad():
        net(x, mods)
    assert ops.launch_count() - n0 > 50
    with open('/proc/self/maps') as f:
        assert 'libhrfuser_b200.so' in f.read()


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_stream_concurrency_and_graph_replay_are_bit_identical(built_lib, precision):
    """forked-stream execution and CUDA-graph replay must not change a single bit"""
    from hrfuser_b200.engine import GraphedForward
    net = _build(backbone_cfg('t', 'nus'), precision).cuda()
    x, mods = synthetic_inputs(2, 128, 160, (3, 3), device='cuda')
    eng = net.engine()
    with torch.no_grad():
        eng.concurrent = False
        a = [t.clone() for t in eng.forward(x, mods)]
        eng.concurrent = True
        b = [t.clone() for t in eng.forward(x, mods)]
        g = GraphedForward(eng, x, mods)
        c = [t.clone() for t in g()]
        c2 = [t.clone() for t in g()]
    torch.cuda.synchronize()
    for p, q, r, r2 in zip(a, b, c, c2):
        assert torch.equal(p, q) and torch.equal(p, r) and torch.equal(p, r2)


STAGE_TAPS = ('fusion_a', 'stage2', 'stage_b', 'fusion_b', 'stage3', 'stage_c', 'fusion_c', 'stage4')


@pytest.mark.parametrize('tag,v,d,mc,H,W', [('t_nus', 't', 'nus', (3, 3), 192, 320),
                                            ('t_stf', 't', 'stf', (3, 2, 1), 96, 320),
                                            ('b_nus', 'b', 'nus', (3, 3), 64, 96)])
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_per_stage_feature_maps(built_lib, tag, v, d, mc, H, W, precision):
    """north_star: per-stage backbone feature maps match the reference -- every fusion-block
    and HRFormer-stage output of the camera stream and of the modality streams, against the
    oracle's taps (themselves pinned on the reference's hooks: tests/test_oracle_golden.py)."""
    cfg = backbone_cfg(v, d)
    net = _build(cfg, precision)
    x, mods = synthetic_inputs(1, H, W, mc, seed=4)
    ref_taps = {}
    with torch.no_grad():
        O.backbone_forward(net.state_dict(), cfg, x, mods, ref_taps)
    net.cuda()
    got_taps = {}
    net.engine().forward(x.cuda(), [m.cuda() for m in mods], taps=got_taps)
    torch.cuda.synchronize()
    n = 0
    for name in STAGE_TAPS:
        assert len(got_taps[name]) == len(ref_taps[name]), name
        for i, (g, r) in enumerate(zip(got_taps[name], ref_taps[name])):
            # bf16 mode: norm-wise <= 2e-2 on every tap; the elementwise share is 97 % here (the
            # small low-resolution maps after ~20 stacked bf16 blocks have a heavier error tail:
            # measured worst case 1.3 % of the elements outside rtol 2e-2) and 99 % on the outputs
            assert_parity(g, r, precision, f'{tag} {name}.{i}', min_frac=0.97)
            n += 1
    M = len(mc)
    assert n == 2 + 2 + M + 3 + 3 + M + 4 + 4


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_t_stf_full_size(built_lib, precision):
    """configs[2] at its shipped size: HRFuser-T Seeing Through Fog, 384 x 1248, cam + lidar +
    radar + gated (3 / 2 / 1 channels)."""
    got, ref = _run(backbone_cfg('t', 'stf'), (3, 2, 1), 1, 384, 1248, precision)
    shapes = [(1, 18, 96, 312), (1, 36, 48, 156), (1, 72, 24, 78), (1, 144, 12, 39)]
    for i, (g, r) in enumerate(zip(got, ref)):
        assert tuple(g.shape) == shapes[i]
        # bf16 mode: norm-wise <= 2e-2 (measured 8.4e-3 on out0); with three modality streams
        # accumulating into the camera stream 98.6 % of out0's elements sit inside rtol 2e-2
        # (T-nus: > 99 %), so the elementwise share asserted here is 98 %
        assert_parity(g, r, precision, f'T-stf 384x1248 out{i}', min_frac=0.98)


def test_b_nus_full_size_bf16(built_lib):
    """configs[3]'s model at its shipped size: HRFuser-B nuScenes 384 x 640, bf16 mode."""
    got, ref = _run(backbone_cfg('b', 'nus'), (3, 3), 1, 384, 640, 'bf16')
    for i, (g, r) in enumerate(zip(got, ref)):
        assert_parity(g, r, 'bf16', f'B-nus 384x640 out{i}')


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_public_forward_graph_cache(built_lib, precision):
    """HRFuserHRFormerBased.forward: first call eager, second call captures a CUDA graph, later
    calls replay it -- all bit-identical, with device and with pinned host inputs, and the
    returned maps are the caller's own (not overwritten by the next call)."""
    net = _build(tiny_cfg(2), precision).cuda()
    x, mods = synthetic_inputs(2, 64, 96, (3, 3), seed=5, device='cuda')
    x2, mods2 = synthetic_inputs(2, 64, 96, (3, 3), seed=6)
    with torch.no_grad():
        a = net(x, mods)                                    # eager
        b = net(x, mods)                                    # capture + replay
        c = net(x, mods)                                    # replay
        d = net(x2.pin_memory(), [m.pin_memory() for m in mods2])      # host inputs, other data
        e = net.engine().forward(x2.cuda(), [m.cuda() for m in mods2])
        torch.cuda.synchronize()
        assert any(v != 0 for v in net._graphs.values())
        for p, q, r in zip(a, b, c):
            assert torch.equal(p, q) and torch.equal(p, r)
        for p, q in zip(d, e):
            assert torch.equal(p, q)
        assert not torch.equal(a[0], d[0])                  # c was not clobbered by the 4th call
        assert torch.equal(a[0], c[0])
        net.load_state_dict(_build(tiny_cfg(2), precision, seed=9).state_dict())
        assert net._graphs == {}                            # a reload drops the captured graphs


def test_parent_load_state_dict_invalidates(built_lib):
    """a checkpoint loaded through a PARENT module never calls the child's
    load_state_dict; the packed weights must be dropped all the same."""
    import torch.nn as nn
    det = nn.Module()
    det.backbone = _build(tiny_cfg(2), 'fp32', seed=1).cuda()
    other = nn.Module()
    other.backbone = _build(tiny_cfg(2), 'fp32', seed=2)
    x, mods = synthetic_inputs(1, 32, 32, (3, 3), device='cuda')
    with torch.no_grad():
        a = det.backbone(x, mods)
        det.load_state_dict(other.state_dict())
        b = det.backbone(x, mods)
        ref = O.backbone_forward(other.backbone.state_dict(), tiny_cfg(2), x.cpu(), [m.cpu() for m in mods])
    assert not torch.equal(a[0], b[0])
    for i, (g, r) in enumerate(zip(b, ref)):
        assert_parity(g, r, 'fp32', f'parent-reloaded out{i}')
