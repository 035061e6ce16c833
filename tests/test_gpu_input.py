"""GPU parity of the input prologue (`hrf_input_prologue_fwd`) through the C-ABI: bit-exact against
the reference-made goldens, against the numpy oracle at the full nuScenes / STF frame sizes, and
through the properties the op has at any size (pad border, channel reversal, zero modality)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import input_oracle

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'input.npz'))
CASES = sorted({k.split('.')[0] for k in GOLD.files})
NUS_IMG = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)
NUS_LIDAR = dict(mean=[0.23277158, 0.31501067, -0.00012928071],
                 std=[2.5538357826888602, 3.7345728854535643, 0.2815488539921788], to_rgb=False)
STF_GATED = dict(mean=[181.74427536], std=[185.49071888], to_rgb=False)
STF_RADAR = dict(mean=[3.4423912, 0.021001821], std=[19.330362993097626, 0.7612592077132296], to_rgb=False)


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize('name', CASES)
def test_bit_exact_against_reference_golden(built_lib, name):
    from hrfuser_b200 import ops
    frames = torch.from_numpy(GOLD[name + '.frames']).cuda()
    got = ops.input_prologue(frames, GOLD[name + '.mean'], GOLD[name + '.std'],
                             to_rgb=bool(GOLD[name + '.to_rgb']), size_divisor=32)
    want = GOLD[name + '.out']
    assert got.dtype == torch.float32 and tuple(got.shape) == want.shape
    assert np.array_equal(bits(got.cpu().numpy()), bits(want))


@pytest.mark.parametrize('shape,dtype,cfg', [
    ((8, 360, 640, 3), np.uint8, NUS_IMG),        # cfg2: 8 frames of 640x360 -> 384x640
    ((8, 360, 640, 3), np.float32, NUS_LIDAR),
    ((2, 384, 1248, 3), np.uint8, NUS_IMG),       # cfg3: STF crop
    ((2, 384, 1248, 1), np.uint8, STF_GATED),
    ((2, 384, 1248, 2), np.float32, STF_RADAR),
    ((1, 666, 1248, 3), np.uint8, NUS_IMG),       # BASELINE.json's literal 1248x666 -> 672x1248
    ((3, 37, 61, 3), np.float32, NUS_LIDAR),      # ragged: W % 4 != 0, unaligned rows
    ((1, 1, 1, 1), np.uint8, STF_GATED),
    ((1, 5, 3, 4), np.float32, dict(mean=[1, 2, 3, 4], std=[2, 3, 4, 5], to_rgb=False)),
])
def test_bit_exact_against_oracle_full_sizes(built_lib, shape, dtype, cfg):
    from hrfuser_b200 import ops
    rng = np.random.default_rng(hash(shape) % 2**32)
    if dtype == np.uint8:
        frames = rng.integers(0, 256, shape, dtype=np.uint8)
    else:
        frames = rng.normal(0, 30, shape).astype(np.float32)
        frames[rng.random(shape) < 0.6] = 0
    want = input_oracle.input_prologue(frames, cfg['mean'], cfg['std'], to_rgb=cfg['to_rgb'], size_divisor=32)
    got = ops.input_prologue(torch.from_numpy(frames).cuda(), cfg['mean'], cfg['std'], to_rgb=cfg['to_rgb'])
    assert tuple(got.shape) == want.shape
    assert np.array_equal(bits(got.cpu().numpy()), bits(want))


def test_properties_at_full_size(built_lib):
    """size-independent checks: pad border, crop-invariance, BGR->RGB == channel flip of no-swap"""
    from hrfuser_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(1)
    frames = torch.randint(0, 256, (8, 360, 640, 3), dtype=torch.uint8, device='cuda', generator=g)
    y = ops.input_prologue(frames, NUS_IMG['mean'], NUS_IMG['std'], to_rgb=True, pad_val=0.0)
    assert tuple(y.shape) == (8, 3, 384, 640)
    assert (y[:, :, 360:, :] == 0).all()
    y7 = ops.input_prologue(frames, NUS_IMG['mean'], NUS_IMG['std'], to_rgb=True, pad_val=7.0)
    assert (y7[:, :, 360:, :] == 7).all() and torch.equal(y7[:, :, :360], y[:, :, :360])
    # swapping the channels of the source and of mean / std instead of to_rgb gives the same tensor
    y2 = ops.input_prologue(frames.flip(-1).contiguous(), NUS_IMG['mean'], NUS_IMG['std'], to_rgb=False)
    assert torch.equal(y2, y)
    # no padding requested: same interior
    y3 = ops.input_prologue(frames, NUS_IMG['mean'], NUS_IMG['std'], to_rgb=True, size_divisor=4)
    assert tuple(y3.shape) == (8, 3, 360, 640) and torch.equal(y3, y[:, :, :360])
    # a frame is independent of its batch neighbours
    y4 = ops.input_prologue(frames[3:4].contiguous(), NUS_IMG['mean'], NUS_IMG['std'], to_rgb=True)
    assert torch.equal(y4[0], y[3])


def test_pipeline_mirror_feeds_the_backbone_shapes(built_lib):
    from hrfuser_b200 import pipeline
    cfgs = [pipeline.Normalize(**NUS_LIDAR, keys=['lidar_img'], sensor_type='lidar'),
            pipeline.Normalize(**NUS_LIDAR, keys=['radar_img'], sensor_type='radar'),
            pipeline.Normalize(**NUS_IMG, keys=['img'])]
    rng = np.random.default_rng(3)
    res = {'img': rng.integers(0, 256, (2, 90, 160, 3), dtype=np.uint8),
           'lidar_img': rng.normal(size=(2, 90, 160, 3)).astype(np.float32),
           'radar_img': np.zeros((2, 90, 160, 3), np.float32)}
    raw = {k: v.copy() for k, v in res.items()}
    out = pipeline.InputPrologue(cfgs, size_divisor=32)(res)
    for k, n in zip(('lidar_img', 'radar_img', 'img'), cfgs):
        want = input_oracle.input_prologue(raw[k], n.mean, n.std, to_rgb=n.to_rgb)
        assert np.array_equal(bits(out[k].cpu().numpy()), bits(want))
    assert out['pad_shape'] == (96, 160, 3) and out['pad_size_divisor'] == 32
    assert set(out) >= {'img_norm_cfg', 'lidar_norm_cfg', 'radar_norm_cfg'}


def test_error_codes(built_lib):
    from hrfuser_b200 import _lib, ops
    x = torch.zeros(1, 8, 8, 3, dtype=torch.uint8, device='cuda')
    with pytest.raises(_lib.HrfError, match='std'):
        ops.input_prologue(x, [0, 0, 0], [1, 0, 1])
    with pytest.raises(_lib.HrfError, match='to_rgb'):
        ops.input_prologue(torch.zeros(1, 8, 8, 2, device='cuda'), [0, 0], [1, 1], to_rgb=True)
    with pytest.raises(_lib.HrfError, match='channels'):
        ops.input_prologue(torch.zeros(1, 8, 8, 5, device='cuda'), [0] * 5, [1] * 5)
    d = _lib.InputDesc(1, 8, 8, 3, 8, 6, _lib.HRF_U8, 0, 0.0)
    m = (C.c_float * 3)(0, 0, 0)
    s = (C.c_float * 3)(1, 1, 1)
    assert built_lib.hrf_input_prologue_fwd(C.byref(d), x.data_ptr(), m, s, x.data_ptr(), None) == -1
