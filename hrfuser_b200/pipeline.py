"""Device-side mirror of the tail of the reference's data pipeline for the backbone's inputs.

The reference prepares every sensor stream on the CPU, per frame, with numpy / OpenCV
(`configs/_base_/datasets/nuscenes_detection_r640_clr_fusion.py:19-31`,
`kitti_detection_2d_c1248_clrg_fusion.py:14-30`):

    Normalize(**lidar_norm_cfg, keys=['lidar_img'], sensor_type='lidar')   transforms.py:719-744
    Normalize(**radar_norm_cfg, keys=['radar_img'], sensor_type='radar')
    Normalize(**img_norm_cfg, keys=['img'])
    Pad(size_divisor=32)                                                    transforms.py:652-667
    DefaultFormatBundle(sensor_keys=[...])                                  formating.py:211-227
    (collate: stack)

`InputPrologue` does those four steps for a batch of raw frames with one
`hrf_input_prologue_fwd` launch per sensor key, so the host ships the raw frames (uint8
camera images: a quarter of the fp32 bytes) and the fp32 NCHW tensors the backbone takes are
written once, on the device.  Same argument names and meaning as the reference's transforms;
like them it adds `<sensor>_norm_cfg`, `pad_shape`, `pad_size_divisor` to the results.

No fallback: without the built library or a CUDA device the call raises (`_lib.HrfError`).
"""
import numpy as np
import torch

from . import ops

SENSOR_TYPES = ('img', 'lidar', 'radar', 'gated')


class Normalize:
    """Argument-compatible with the reference's `Normalize` (transforms.py:706-717)."""

    def __init__(self, mean, std, to_rgb=True, keys=None, with_mask=None, sensor_type='img'):
        self.mean = np.array(mean, dtype=np.float32)
        self.std = np.array(std, dtype=np.float32)
        self.to_rgb = to_rgb
        self.keys = keys
        if with_mask:
            raise NotImplementedError('with_mask is not used by any shipped config')
        self.sensor_type = sensor_type
        if self.sensor_type not in SENSOR_TYPES:
            raise Exception('Sensor type not supported')


class InputPrologue:
    """`InputPrologue([Normalize(...), ...], size_divisor=32)(results) -> results`.

    `results[key]` is a batch of raw frames: a `(B, H, W, C)` (or `(B, H, W)` for one channel)
    uint8 / float32 torch tensor or numpy array, on the host or already on the device.  On
    return `results[key]` is the fp32 `(B, C, Hp, Wp)` CUDA tensor
    `HRFuserHRFormerBased.forward` takes.
    """

    def __init__(self, normalizers, size_divisor=32, pad_val=0, device='cuda'):
        self.normalizers = list(normalizers)
        self.size_divisor = size_divisor
        self.pad_val = pad_val.get('img', 0) if isinstance(pad_val, dict) else pad_val
        self.device = torch.device(device)

    def _frames(self, a):
        t = torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
        if t.dim() == 3:
            t = t.unsqueeze(-1)
        if t.dim() != 4:
            raise ValueError('sensor frames must be (B,H,W,C) or (B,H,W)')
        if t.dtype not in (torch.uint8, torch.float32):
            t = t.to(torch.float32)          # DefaultFormatBundle / imnormalize: astype(float32)
        return t.to(self.device, non_blocking=True).contiguous()

    def __call__(self, results):
        padded = None
        for n in self.normalizers:
            for key in (n.keys or ['img']):
                x = ops.input_prologue(self._frames(results[key]), n.mean, n.std, to_rgb=n.to_rgb,
                                       size_divisor=self.size_divisor, pad_val=self.pad_val)
                results[key] = x
                padded = x
            results[n.sensor_type + '_norm_cfg'] = dict(mean=n.mean, std=n.std, to_rgb=n.to_rgb)
        if padded is not None:
            results['pad_shape'] = (padded.shape[2], padded.shape[3], padded.shape[1])
            results['pad_fixed_size'] = None
            results['pad_size_divisor'] = self.size_divisor
        return results
