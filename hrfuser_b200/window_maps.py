"""Integer geometry of the window pad / partition / merge step.

Closed forms of what the reference does with `F.pad` + `view/permute/reshape`
(reference hrformer.py:196-209,229-236; hrfuser_hrformer_based.py:201-221,
241-248) and of the relative-position index it builds with meshgrid
(hrformer.py:63-82; hrfuser_hrformer_based.py:71-90).  The CUDA kernels use the
same closed forms for their addressing (csrc/window_attn.cuh), so no window
tensor is ever materialised; `tests/test_window_maps.py` checks these maps
bit-exactly against the oracle and the committed golden maps.
"""
from collections import namedtuple

import numpy as np

WindowGeometry = namedtuple(
    'WindowGeometry', 'H W Wh Ww nWh nWw Hp Wp pad_h pad_w pad_t pad_b pad_l pad_r')


def window_geometry(H, W, Wh=7, Ww=7):
    nWh, nWw = -(-H // Wh), -(-W // Ww)
    pad_h, pad_w = nWh * Wh - H, nWw * Ww - W
    return WindowGeometry(H, W, Wh, Ww, nWh, nWw, nWh * Wh, nWw * Ww, pad_h, pad_w,
                          pad_h // 2, pad_h - pad_h // 2, pad_w // 2, pad_w - pad_w // 2)


def relative_position_index(Wh=7, Ww=7):
    """(Wh*Ww, Wh*Ww) int64: index into the (2Wh-1)(2Ww-1)-row bias table for
    query slot i and key slot j:  (ih-jh+Wh-1)*(2Ww-1) + (iw-jw+Ww-1)."""
    s = np.arange(Wh * Ww)
    h, w = s // Ww, s % Ww
    return ((h[:, None] - h[None, :] + Wh - 1) * (2 * Ww - 1)
            + (w[:, None] - w[None, :] + Ww - 1)).astype(np.int64)


def token_to_window(H, W, Wh=7, Ww=7):
    """For every real token (h, w): (window id within the image, slot inside the
    window), both int32 arrays of shape (H, W)."""
    g = window_geometry(H, W, Wh, Ww)
    hp = np.arange(H)[:, None] + g.pad_t
    wp = np.arange(W)[None, :] + g.pad_l
    win = (hp // Wh) * g.nWw + (wp // Ww)
    slot = (hp % Wh) * Ww + (wp % Ww)
    return win.astype(np.int32), slot.astype(np.int32)


def window_to_token(H, W, Wh=7, Ww=7):
    """Inverse map, shape (nW, Wh*Ww) int32: flat token index h*W+w held by each
    window slot, or -1 where the slot is zero padding."""
    g = window_geometry(H, W, Wh, Ww)
    out = np.full((g.nWh * g.nWw, Wh * Ww), -1, np.int32)
    win, slot = token_to_window(H, W, Wh, Ww)
    out[win.reshape(-1), slot.reshape(-1)] = np.arange(H * W, dtype=np.int32)
    return out
