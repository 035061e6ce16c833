"""Generate the committed golden vectors by RUNNING THE UNMODIFIED REFERENCE.

    python tests/golden/make_golden.py            (needs /root/reference or $HRFUSER_REF)

The reference repository ships no tests or fixtures for the backbone (SURVEY.md
section 4), so these files -- outputs of its own modules on seeded inputs with
seeded, fully randomised parameters -- are the anchor that pins
oracle/hrfuser_oracle.py and, through it, the CUDA path.  Parameters are not
stored (HRFuser-T has 4 M): `hrfuser_b200.utils.randomize_parameters(module,
seed)` draws them in state_dict order, which is identical for the reference
module and the drop-in; `param_checksum` in every file detects any drift of the
generator.  Everything here is test infrastructure.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from hrfuser_b200.configs import backbone_cfg  # noqa: E402
from hrfuser_b200.utils import randomize_parameters, synthetic_inputs  # noqa: E402
from oracle import ref_loader  # noqa: E402

BN = dict(type='BN', requires_grad=True, momentum=0.1)
LN = dict(type='LN', eps=1e-6)
GRIDS = [(96, 160), (48, 80), (24, 40), (12, 20), (96, 312), (48, 156), (24, 78), (12, 39),
         (7, 7), (14, 21), (5, 3), (1, 1)]


def checksum(module):
    return float(sum(v.double().sum() for v in module.state_dict().values() if v.is_floating_point()))


def layout_digest(sd):
    txt = '\n'.join(f'{k} {tuple(v.shape)} {v.dtype}' for k, v in sd.items())
    return hashlib.sha256(txt.encode()).hexdigest()


def golden_maps(ref, hrformer):
    out = {}
    lsa = hrformer.LocalWindowSelfAttention(2, 1, 7)
    out['relative_position_index'] = lsa.attn.relative_position_index.numpy().astype(np.int64)
    captured = {}

    class Capture(torch.nn.Module):
        def forward(self, x, *a, **k):
            captured['w'] = x.clone()
            return x
    lsa.attn = Capture()
    for (H, W) in GRIDS:
        ids = torch.arange(1, H * W + 1, dtype=torch.float64).view(1, H * W, 1)   # 0 == padding
        back = lsa(ids, H, W)
        assert torch.equal(back, ids)                         # merge + crop invert the partition
        out[f'window_to_token_{H}x{W}'] = (captured['w'][..., 0].long() - 1).numpy().astype(np.int64)
    return out


def golden_ops(ref, hrformer):
    out = {}
    torch.manual_seed(0)
    # LSA block / FFN (HRFormerBlock = x + LSA(LN x); x + FFN(LN x))
    for name, (C, heads, H, W, B) in {'blk_c18': (18, 1, 12, 20, 2), 'blk_c36': (36, 2, 9, 13, 1),
                                      'blk_c72': (72, 4, 7, 7, 1)}.items():
        blk = hrformer.HRFormerBlock(C, C, heads, 7, 4, 0., norm_cfg=BN, transformer_norm_cfg=LN)
        randomize_parameters(blk, 11)
        blk.eval()
        x = torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(C)) * 3
        with torch.no_grad():
            t = x.flatten(2).transpose(1, 2)
            after_attn = t + blk.attn(blk.norm1(t), H, W)
            y = blk(x)
        out[name + '_x'] = x.numpy()
        out[name + '_after_attn'] = after_attn.numpy()
        out[name + '_y'] = y.numpy()
        out[name + '_cfg'] = np.array([C, heads, H, W, B])
        out[name + '_checksum'] = np.array(checksum(blk))
    # fusion block (MWCA), M = 2 and 3
    for name, (C, heads, H, W, M) in {'fus_c36_m2': (36, 2, 12, 20, 2), 'fus_c18_m3': (18, 1, 9, 10, 3)}.items():
        blk = ref.HRFuserFusionBlock(C, C, heads, 7, 4, 0., norm_cfg=BN, transformer_norm_cfg=LN,
                                     num_fused_modalities=M, proj_drop_rate=0.1)
        randomize_parameters(blk, 12)
        blk.eval()
        g = torch.Generator().manual_seed(C + M)
        x = torch.randn(1, C, H, W, generator=g) * 3
        zs = [torch.randn(1, C, H, W, generator=g) * 3 for _ in range(M)]
        with torch.no_grad():
            y = blk(x, zs)
        out[name + '_x'] = x.numpy()
        for k, z in enumerate(zs):
            out[f'{name}_z{k}'] = z.numpy()
        out[name + '_y'] = y.numpy()
        out[name + '_cfg'] = np.array([C, heads, H, W, M])
        out[name + '_checksum'] = np.array(checksum(blk))
    # multi-resolution exchange: a 3-branch HRFomerModule with no blocks' effect removed
    # (blocks run too; the golden is the whole module)
    mod = hrformer.HRFomerModule(3, ref.HRFormerBlock, (1, 1, 1), [18, 36, 72], [18, 36, 72],
                                 (1, 2, 4), (7, 7, 7), (4, 4, 4), True, drop_paths=[0.0],
                                 norm_cfg=BN, transformer_norm_cfg=LN)
    randomize_parameters(mod, 13)
    mod.eval()
    g = torch.Generator().manual_seed(5)
    xs = [torch.randn(1, c, 16 >> i, 24 >> i, generator=g) * 3 for i, c in enumerate((18, 36, 72))]
    with torch.no_grad():
        ys = mod([x.clone() for x in xs])
    for i in range(3):
        out[f'mod3_x{i}'] = xs[i].numpy()
        out[f'mod3_y{i}'] = ys[i].numpy()
    out['mod3_checksum'] = np.array(checksum(mod))
    return out


def golden_e2e():
    out, layout = {}, {}
    for tag, (v, d, mc) in {'t_nus': ('t', 'nus', (3, 3)), 't_stf': ('t', 'stf', (3, 2, 1)),
                            'b_nus': ('b', 'nus', (3, 3))}.items():
        cfg = backbone_cfg(v, d)
        net = ref_loader.build_reference_backbone(cfg)
        randomize_parameters(net, 1)
        sd = net.state_dict()
        layout[tag] = dict(count=len(sd), digest=layout_digest(sd),
                           params=sum(p.numel() for p in net.parameters()))
        H, W = (64, 96) if tag != 'b_nus' else (64, 64)
        x, mods = synthetic_inputs(1, H, W, mc, seed=3)
        taps = {}
        hooks = []
        for name in ('fusion_a', 'fusion_b', 'fusion_c'):
            for i, m in enumerate(getattr(net, name)):
                hooks.append(m.register_forward_hook(
                    lambda mod, a, o, key=f'{name}.{i}': taps.__setitem__(key, o.detach().clone())))
        for name in ('stage2', 'stage3', 'stage4'):
            hooks.append(getattr(net, name).register_forward_hook(
                lambda mod, a, o, key=name: taps.__setitem__(key, [t.detach().clone() for t in o])))
        with torch.no_grad():
            ys = net(x, [m.clone() for m in mods])
        for h in hooks:
            h.remove()
        out[f'{tag}_hw'] = np.array([H, W])
        out[f'{tag}_checksum'] = np.array(checksum(net))
        for i, y in enumerate(ys):
            out[f'{tag}_out{i}'] = y.numpy()
        for k, v_ in taps.items():
            if isinstance(v_, list):
                for i, t in enumerate(v_):
                    out[f'{tag}_{k}.{i}'] = t.numpy()
            else:
                out[f'{tag}_{k}'] = v_.numpy()
        if tag == 't_nus':
            # configs[0] at full size: statistics + a fixed sample of every output
            x, mods = synthetic_inputs(1, 384, 640, mc, seed=0)
            with torch.no_grad():
                ys = net(x, [m.clone() for m in mods])
            for i, y in enumerate(ys):
                yd = y.double()
                out[f'full_t_nus_out{i}_chan_mean'] = yd.mean((0, 2, 3)).numpy()
                out[f'full_t_nus_out{i}_chan_sqmean'] = (yd ** 2).mean((0, 2, 3)).numpy()
                idx = torch.randperm(y.numel(), generator=torch.Generator().manual_seed(i))[:4096]
                out[f'full_t_nus_out{i}_idx'] = idx.numpy()
                out[f'full_t_nus_out{i}_val'] = y.flatten()[idx].numpy()
    return out, layout


def main():
    root = ref_loader.find_reference()
    if root is None:
        sys.exit('no reference checkout found')
    ref = ref_loader.load_reference(root)
    hrformer = sys.modules['mmdet.models.backbones.hrformer']
    np.savez_compressed(os.path.join(HERE, 'maps.npz'), **golden_maps(ref, hrformer))
    np.savez_compressed(os.path.join(HERE, 'ops.npz'), **golden_ops(ref, hrformer))
    e2e, layout = golden_e2e()
    np.savez_compressed(os.path.join(HERE, 'e2e.npz'), **e2e)
    with open(os.path.join(HERE, 'state_dict_layout.json'), 'w') as f:
        json.dump(dict(torch=torch.__version__, layouts=layout), f, indent=1)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == '__main__':
    main()
