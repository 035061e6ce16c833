"""Window index / padding maps: product closed forms == oracle == reference golden, bit-exact."""
import os

import numpy as np
import pytest

from hrfuser_b200 import window_maps as wm
from oracle import hrfuser_oracle as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'maps.npz'))
GRIDS = sorted(tuple(int(v) for v in k.split('_')[-1].split('x'))
               for k in GOLD.files if k.startswith('window_to_token_'))


def test_relative_position_index_bit_exact():
    g = GOLD['relative_position_index']
    assert g.dtype == np.int64 and g.shape == (49, 49)
    assert np.array_equal(wm.relative_position_index(7, 7), g)
    assert np.array_equal(O.relative_position_index(7, 7), g)
    # closed form of SURVEY.md section 8c
    i, j = np.divmod(np.arange(49), 7), np.divmod(np.arange(49), 7)
    assert np.array_equal(g, (i[0][:, None] - j[0][None] + 6) * 13 + (i[1][:, None] - j[1][None] + 6))
    assert g.min() == 0 and g.max() == 168


@pytest.mark.parametrize('H,W', GRIDS)
def test_window_to_token_bit_exact(H, W):
    g = GOLD[f'window_to_token_{H}x{W}']
    assert np.array_equal(wm.window_to_token(H, W), g)
    assert np.array_equal(O.window_gather_map(H, W), g)


@pytest.mark.parametrize('H,W', GRIDS)
def test_partition_is_a_bijection_on_real_tokens(H, W):
    win, slot = wm.token_to_window(H, W)
    inv = wm.window_to_token(H, W)
    assert np.array_equal(inv[win.reshape(-1), slot.reshape(-1)], np.arange(H * W))
    assert (inv >= 0).sum() == H * W            # every other slot is padding
    geo = wm.window_geometry(H, W)
    assert geo.Hp % 7 == 0 and geo.Wp % 7 == 0 and 0 <= geo.pad_h < 7 and 0 <= geo.pad_w < 7
    assert (geo.pad_t, geo.pad_b, geo.pad_l, geo.pad_r) == O.pad_amounts(H, W)
    assert inv.shape == (geo.nWh * geo.nWw, 49)


def test_paddings_of_the_shipped_grids():
    # SURVEY.md App. A item 2
    want = {(96, 160): (98, 161), (48, 80): (49, 84), (24, 40): (28, 42), (12, 20): (14, 21),
            (96, 312): (98, 315), (48, 156): (49, 161), (24, 78): (28, 84), (12, 39): (14, 42)}
    for (H, W), (Hp, Wp) in want.items():
        g = wm.window_geometry(H, W)
        assert (g.Hp, g.Wp) == (Hp, Wp)


@pytest.mark.parametrize('H,W', [(168, 312), (84, 156), (42, 78), (21, 39)])
def test_full_stf_frame_grids_product_equals_oracle(H, W):
    """workload 'hrfuser_t_stf_r1248_full' (BASELINE.json's literal 1248x666 -> 672x1248): no
    reference golden for these grids, so the product closed form is held to the oracle's
    gather map (itself pinned on the eight shipped grids above) and to the bijection."""
    inv = wm.window_to_token(H, W)
    assert np.array_equal(inv, O.window_gather_map(H, W))
    win, slot = wm.token_to_window(H, W)
    assert np.array_equal(inv[win.reshape(-1), slot.reshape(-1)], np.arange(H * W))
    assert (inv >= 0).sum() == H * W
