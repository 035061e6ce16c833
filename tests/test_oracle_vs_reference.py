"""Pin the oracle against the live, unmodified reference (build container only)."""
import numpy as np
import pytest
import torch

from hrfuser_b200 import backbone_cfg
from hrfuser_b200.utils import randomize_parameters, rel_err, synthetic_inputs
from oracle import hrfuser_oracle as O
from oracle import ref_loader

pytestmark = pytest.mark.skipif(ref_loader.find_reference() is None,
                                reason='no reference checkout (expected on the GPU box)')


@pytest.mark.parametrize('v,d,mc,hw', [('t', 'nus', (3, 3), (96, 160)), ('t', 'stf', (3, 2, 1), (64, 128)),
                                       ('b', 'nus', (3, 3), (64, 96))])
def test_backbone_matches_reference(v, d, mc, hw):
    cfg = backbone_cfg(v, d)
    ref = ref_loader.build_reference_backbone(cfg)
    randomize_parameters(ref, 7)
    x, mods = synthetic_inputs(2, *hw, mc, seed=9, sparse=True)
    with torch.no_grad():
        want = ref(x, [m.clone() for m in mods])
        got = O.backbone_forward(ref.state_dict(), cfg, x, mods)
    for g, w in zip(got, want):
        assert rel_err(g, w) < 2e-6


def test_pad_mask_variant_matches_reference():
    cfg = backbone_cfg('t', 'nus')
    cfg['extra']['with_pad_mask'] = True
    ref = ref_loader.build_reference_backbone(cfg)
    randomize_parameters(ref, 8)
    x, mods = synthetic_inputs(1, 96, 160, (3, 3), seed=2)
    with torch.no_grad():
        want = ref(x, [m.clone() for m in mods])
        got = O.backbone_forward(ref.state_dict(), cfg, x, mods)
    for g, w in zip(got, want):
        assert rel_err(g, w) < 2e-6


def test_reference_configs_resolve_to_our_dicts():
    for v, d, f in [('t', 'nus', 'cascade_rcnn_hrfuser_t_1x_nus_r640_l_r_fusion_bn.py'),
                    ('t', 'stf', 'cascade_rcnn_hrfuser_t_1x_stf_r1248_4mod_bn.py'),
                    ('b', 'nus', 'cascade_rcnn_hrfuser_b_1x_nus_r640_l_r_fusion_bn.py')]:
        rc = ref_loader.read_reference_config('configs/hrfuser/' + f)['model']['backbone']
        assert rc == backbone_cfg(v, d)
    rc = ref_loader.read_reference_config(
        'configs/hrfuser/cascade_rcnn_hrfuser_t_1x_nus_r640_l_r_fusion.py')['model']['backbone']
    assert rc == backbone_cfg('t', 'nus', norm='SyncBN')


def test_state_dict_layout_matches_reference():
    import copy
    from hrfuser_b200 import HRFuserHRFormerBased
    for v, d in [('t', 'nus'), ('t', 'stf'), ('b', 'nus')]:
        cfg = backbone_cfg(v, d)
        ref = ref_loader.build_reference_backbone(cfg)
        c = copy.deepcopy(cfg)
        c.pop('type')
        mine = HRFuserHRFormerBased(**c)
        a, b = ref.state_dict(), mine.state_dict()
        assert list(a) == list(b)
        assert all(a[k].shape == b[k].shape and a[k].dtype == b[k].dtype for k in a)
        idx = [k for k in a if k.endswith('relative_position_index')]
        assert idx and all(torch.equal(a[k], b[k]) for k in idx)
