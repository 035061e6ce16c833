"""Input prologue, CPU side: the numpy oracle against the golden vectors made
from the reference's own Normalize / Pad / DefaultFormatBundle classes
(tests/golden/make_golden_input.py), and the host-side mirror's argument handling."""
import os

import numpy as np
import pytest
import torch

from oracle import input_oracle
from hrfuser_b200 import _lib, pipeline

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'input.npz'))
CASES = sorted({k.split('.')[0] for k in GOLD.files})


def case(name):
    return (GOLD[name + '.frames'], GOLD[name + '.mean'], GOLD[name + '.std'], bool(GOLD[name + '.to_rgb']),
            GOLD[name + '.out'])


def test_golden_has_the_expected_cases():
    assert {'nus_img', 'nus_lidar', 'nus_radar', 'stf_gated', 'stf_radar', 'stf_img_f32_ragged',
            'aligned_no_pad', 'zero_modality'} <= set(CASES)


@pytest.mark.parametrize('name', CASES)
def test_oracle_bit_exact_against_reference_golden(name):
    frames, mean, std, to_rgb, want = case(name)
    if frames.shape[-1] == 1:
        frames = frames[..., 0]                      # the reference loads 1-channel images as (H,W)
    got = input_oracle.input_prologue(frames, mean, std, to_rgb=to_rgb, size_divisor=32)
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_pad_region_is_pad_val_not_normalised():
    frames, mean, std, to_rgb, want = case('nus_img')
    H, W = frames.shape[1:3]
    assert (want[:, :, H:, :] == 0).all() and (want[:, :, :, W:] == 0).all()
    got = input_oracle.input_prologue(frames, mean, std, to_rgb=to_rgb, size_divisor=32, pad_val=7)
    assert (got[:, :, H:, :] == 7).all()


def test_zero_modality_is_minus_mean_over_std():
    frames, mean, std, to_rgb, want = case('zero_modality')
    H, W = frames.shape[1:3]
    m32, s32 = mean.astype(np.float32), std.astype(np.float32)
    for c in range(3):
        v = np.float32(np.float64(np.float32(0) - m32[c]) * (1.0 / np.float64(s32[c])))
        assert (want[0, c, :H, :W] == v).all()


def test_normalize_mirror_arguments():
    n = pipeline.Normalize(mean=[1, 2, 3], std=[4, 5, 6], to_rgb=False, keys=['lidar_img'], sensor_type='lidar')
    assert n.mean.dtype == np.float32 and n.std.dtype == np.float32 and n.keys == ['lidar_img']
    with pytest.raises(Exception, match='Sensor type not supported'):     # transforms.py:716-717
        pipeline.Normalize(mean=[0], std=[1], sensor_type='thermal')


def test_prologue_without_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    p = pipeline.InputPrologue([pipeline.Normalize(mean=[0, 0, 0], std=[1, 1, 1], keys=['img'])], device='cpu')
    with pytest.raises((ValueError, _lib.HrfError)):
        p({'img': np.zeros((1, 8, 8, 3), np.uint8)})
