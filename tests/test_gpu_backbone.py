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
            # stacked blocks: norm-wise bound end to end
            assert rel_err(g.cpu(), r) < 3e-2, (i, rel_err(g.cpu(), r))


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
        e = rel_err(g.cpu(), r)
        assert e < 3e-2, f'T-nus bf16 out{i}: norm-wise {e:.3e}'


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
        if precision == 'fp32':
            assert_parity(g, r, 'fp32', f'B-nus out{i}')
        else:
            assert rel_err(g.cpu(), r) < 3e-2, (i, rel_err(g.cpu(), r))


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
