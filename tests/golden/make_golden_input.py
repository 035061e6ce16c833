"""Generates tests/golden/input.npz: raw sensor frames and the output of the reference's own
`Normalize` -> `Pad(size_divisor=32)` -> `DefaultFormatBundle` classes on them
(mmdet/datasets/pipelines/transforms.py:608-756, formating.py:197-227), imported from the
read-only reference checkout.

mmcv (mmcv-full 1.3.17) is not installed; the three image functions the classes call are
supplied by a stand-in that issues the same OpenCV calls mmcv does (cv2 is installed):
`imnormalize` = copy().astype(float32), cvtColor BGR2RGB, cv2.subtract(mean f64),
cv2.multiply(1/f64(std)); `impad` / `impad_to_multiple` = cv2.copyMakeBorder(BORDER_CONSTANT)
on the bottom / right.  `mmdet.core`, `mmdet.datasets.builder` are empty stubs (registry
decorator, unused mask classes).

Run from the repo root:  python tests/golden/make_golden_input.py
"""
import importlib.util
import os
import sys
import types

import cv2
import numpy as np
import torch

REF = os.environ.get('HRFUSER_REF', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))


def _mmcv_standin():
    mmcv = types.ModuleType('mmcv')

    def imnormalize(img, mean, std, to_rgb=True):
        img = img.copy().astype(np.float32)
        assert img.dtype != np.uint8
        mean = np.float64(mean.reshape(1, -1))
        stdinv = 1 / np.float64(std.reshape(1, -1))
        if to_rgb:
            cv2.cvtColor(img, cv2.COLOR_BGR2RGB, img)
        cv2.subtract(img, mean, img)
        cv2.multiply(img, stdinv, img)
        return img

    def impad(img, *, shape=None, padding=None, pad_val=0, padding_mode='constant'):
        assert padding_mode == 'constant' and shape is not None
        pad_h, pad_w = shape[0] - img.shape[0], shape[1] - img.shape[1]
        return cv2.copyMakeBorder(img, 0, pad_h, 0, pad_w, cv2.BORDER_CONSTANT, value=pad_val)

    def impad_to_multiple(img, divisor, pad_val=0):
        pad_h = int(np.ceil(img.shape[0] / divisor)) * divisor
        pad_w = int(np.ceil(img.shape[1] / divisor)) * divisor
        return impad(img, shape=(pad_h, pad_w), pad_val=pad_val)

    mmcv.imnormalize, mmcv.impad, mmcv.impad_to_multiple = imnormalize, impad, impad_to_multiple
    mmcv.is_str = lambda x: isinstance(x, str)
    par = types.ModuleType('mmcv.parallel')

    class DataContainer:
        def __init__(self, data, stack=False, padding_value=0, cpu_only=False, pad_dims=2):
            self.data, self.stack = data, stack

    par.DataContainer = DataContainer
    mmcv.parallel = par
    return mmcv, par


def load_reference_pipeline():
    mmcv, par = _mmcv_standin()
    sys.modules.update({'mmcv': mmcv, 'mmcv.parallel': par})
    for name in ('mmdet', 'mmdet.core', 'mmdet.core.evaluation', 'mmdet.core.evaluation.bbox_overlaps',
                 'mmdet.datasets', 'mmdet.datasets.builder', 'mmdet.datasets.pipelines'):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    sys.modules['mmdet.core'].PolygonMasks = object
    sys.modules['mmdet.core'].find_inside_bboxes = None
    sys.modules['mmdet.core.evaluation.bbox_overlaps'].bbox_overlaps = None

    class _Reg:
        def register_module(self, *a, **k):
            return lambda cls: cls

    sys.modules['mmdet.datasets.builder'].PIPELINES = _Reg()
    mods = {}
    for fn in ('transforms', 'formating'):
        spec = importlib.util.spec_from_file_location(
            f'mmdet.datasets.pipelines.{fn}', os.path.join(REF, 'mmdet', 'datasets', 'pipelines', fn + '.py'))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        mods[fn] = mod
    return mods['transforms'], mods['formating']


# sensor streams of the two shipped dataset configs
# (configs/_base_/datasets/nuscenes_detection_r640_clr_fusion.py:12-17,
#  configs/_base_/datasets/kitti_detection_2d_c1248_clrg_fusion.py:5-12)
CASES = {
    # name: (key, sensor_type, dtype, (B,H,W,C), mean, std, to_rgb)
    'nus_img': ('img', 'img', 'uint8', (2, 90, 160, 3), [123.675, 116.28, 103.53], [58.395, 57.12, 57.375], True),
    'nus_lidar': ('lidar_img', 'lidar', 'float32', (2, 90, 160, 3), [0.23277158, 0.31501067, -0.00012928071],
                  [2.5538357826888602, 3.7345728854535643, 0.2815488539921788], False),
    'nus_radar': ('radar_img', 'radar', 'float32', (2, 90, 160, 3), [0.19778967, 0.03477772, 0.0025186215],
                  [3.219927182957935, 0.7240392925308506, 0.11561270078715341], False),
    'stf_gated': ('gated_img', 'gated', 'uint8', (2, 48, 78, 1), [181.74427536], [185.49071888], False),
    'stf_radar': ('radar_img', 'radar', 'float32', (1, 48, 78, 2), [3.4423912, 0.021001821],
                  [19.330362993097626, 0.7612592077132296], False),
    'stf_img_f32_ragged': ('img', 'img', 'float32', (1, 37, 61, 3), [95.07200648, 91.35659045, 87.7264499],
                           [42.78716034, 42.98587388, 43.82545466], True),
    'aligned_no_pad': ('img', 'img', 'uint8', (1, 64, 96, 3), [123.675, 116.28, 103.53], [58.395, 57.12, 57.375], True),
    'zero_modality': ('lidar_img', 'lidar', 'float32', (1, 33, 40, 3), [0.23277158, 0.31501067, -0.00012928071],
                      [2.5538357826888602, 3.7345728854535643, 0.2815488539921788], False),
}


def make_frames(name, dtype, shape, rng):
    if name == 'zero_modality':          # RandomDrop zeroes a whole modality (transforms.py:487-507)
        return np.zeros(shape, dtype)
    if dtype == 'uint8':
        return rng.integers(0, 256, shape, dtype=np.uint8)
    x = rng.normal(0, 20, shape).astype(np.float32)
    x[rng.random(shape) < 0.7] = 0.0     # projected lidar / radar images are mostly empty
    return x


def main():
    T, F = load_reference_pipeline()
    rng = np.random.default_rng(0)
    out = {}
    for name, (key, sensor, dtype, shape, mean, std, to_rgb) in CASES.items():
        frames = make_frames(name, dtype, shape, rng)
        ys = []
        for f in frames:
            res = {key: f if shape[3] > 1 else f[..., 0], 'img_fields': [key]}
            res = T.Normalize(mean=mean, std=std, to_rgb=to_rgb, keys=[key], sensor_type=sensor)(res)
            res = T.Pad(size_divisor=32)(res)
            if key != 'img':
                res = F.DefaultFormatBundle(sensor_keys=[key])(res)
                y = res[key].data
            else:                         # 'img' also wants meta keys; same code path for the tensor
                res.setdefault('img', res[key])
                res = F.DefaultFormatBundle(sensor_keys=[key])(res)
                y = res[key].data
            ys.append(y)
        y = torch.stack(ys).numpy()
        assert y.dtype == np.float32
        out[name + '.frames'] = frames
        out[name + '.out'] = y
        out[name + '.mean'] = np.array(mean, np.float64)
        out[name + '.std'] = np.array(std, np.float64)
        out[name + '.to_rgb'] = np.array(int(to_rgb))
        print(name, frames.shape, frames.dtype, '->', y.shape)
    np.savez_compressed(os.path.join(HERE, 'input.npz'), **out)


if __name__ == '__main__':
    main()
