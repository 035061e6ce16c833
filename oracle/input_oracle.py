"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the tail of the reference's data
pipeline for one sensor stream: Normalize -> Pad(size_divisor) -> DefaultFormatBundle ->
collate.  Only `tests/` may import this; the product (`hrfuser_b200/`) never does.

Follows
  * `Normalize.__call__`            mmdet/datasets/pipelines/transforms.py:719-744
  * `mmcv.imnormalize`              third-party, mmcv-full 1.3.17 (README.md:41), NOT in the
    reference tree.  Published algorithm: `img.copy().astype(np.float32)`; if `to_rgb`,
    `cv2.cvtColor(img, COLOR_BGR2RGB, img)`; `cv2.subtract(img, np.float64(mean), img)`;
    `cv2.multiply(img, 1 / np.float64(std), img)`.  OpenCV (4.13, measured bit for bit by
    the golden script; an all-fp32 and an all-fp64 restatement both miss by 1 ulp on 3-27 % of
    the pixels) subtracts a scalar from a CV_32F array in fp32 and multiplies by a scalar in
    fp64, so each output is `fl32(f64(fl32(x - mean32)) * (1 / f64(std32)))`.
  * `Pad._pad_img`                  transforms.py:652-667 (`mmcv.impad_to_multiple`: constant
    border on the bottom / right up to the next multiple of `size_divisor`)
  * `DefaultFormatBundle.__call__`  formating.py:211-227 (uint8 -> float32, (H,W) -> (H,W,1),
    HWC -> CHW)

Pinned (tests/golden/input.npz, made by tests/golden/make_golden_input.py): bit-exact against
the reference's own `Normalize` / `Pad` / `DefaultFormatBundle` classes imported from the
reference tree, with `mmcv.imnormalize` / `impad_to_multiple` supplied by a stand-in that makes
the OpenCV (cv2 4.13, installed here) calls listed above.
"""
import numpy as np


def imnormalize(img, mean, std, to_rgb):
    img = np.asarray(img).astype(np.float32)
    if img.ndim == 2:
        img = img[..., None]
    mean = np.asarray(mean, dtype=np.float32).reshape(1, 1, -1)
    stdinv = 1.0 / np.asarray(std, dtype=np.float32).astype(np.float64)
    if to_rgb:
        assert img.shape[-1] == 3
        img = img[..., ::-1]
    diff = (img - mean).astype(np.float32)
    return (diff.astype(np.float64) * stdinv.reshape(1, 1, -1)).astype(np.float32)


def impad_to_multiple(img, divisor, pad_val=0):
    h, w = img.shape[:2]
    hp, wp = -(-h // divisor) * divisor, -(-w // divisor) * divisor
    out = np.full((hp, wp) + img.shape[2:], pad_val, dtype=img.dtype)
    out[:h, :w] = img
    return out


def input_prologue(frames, mean, std, to_rgb=False, size_divisor=32, pad_val=0):
    """frames: (B,H,W,C) or (B,H,W) uint8 / float32 -> (B,C,Hp,Wp) float32"""
    outs = []
    for f in frames:
        y = imnormalize(f, mean, std, to_rgb)
        y = impad_to_multiple(y, size_divisor, pad_val)
        outs.append(np.ascontiguousarray(y.transpose(2, 0, 1)))
    return np.stack(outs)
