"""Small host-side helpers shared by tests, the benchmark and the smoke test."""
import torch


@torch.no_grad()
def randomize_parameters(module, seed=0):
    """Seeded, *non-degenerate* values for every parameter and BN buffer.

    Default construction leaves the relative-position tables at zero and the BN
    statistics at identity, which would hide bugs in exactly the paths this
    repo re-implements (SURVEY.md App. F recipe): weights ~ N(0, 1/fan_in),
    biases ~ N(0, 0.1^2), norm scales 1 + 0.1 N(0,1), BN running_mean
    0.1 N(0,1), running_var 1 + 0.2 U(0,1).  The draw order is the order of
    `named_parameters()` / `named_buffers()`, which is identical for the
    reference module and the drop-in (same state_dict layout), so both get
    the same values from the same seed.
    """
    g = torch.Generator().manual_seed(seed)

    def normal(shape, std):
        return torch.randn(shape, generator=g) * std
    for name, p in module.named_parameters():
        leaf = name.rsplit('.', 1)[-1]
        if name.endswith('relative_position_bias_table'):
            v = normal(p.shape, 0.5)
        elif p.dim() == 1 and leaf == 'weight':           # norm scale
            v = 1.0 + normal(p.shape, 0.1)
        elif leaf == 'bias':
            v = normal(p.shape, 0.1)
        else:                                             # conv / linear weight
            fan_in = p[0].numel()
            v = normal(p.shape, fan_in ** -0.5)
        p.copy_(v.to(p.dtype))
    for name, b in module.named_buffers():
        if name.endswith('running_mean'):
            b.copy_(normal(b.shape, 0.1))
        elif name.endswith('running_var'):
            b.copy_(1.0 + 0.2 * torch.rand(b.shape, generator=g))
    if hasattr(module, 'invalidate_engine'):
        module.invalidate_engine()
    return module


def synthetic_inputs(batch, H, W, mod_channels=(3, 3), seed=0, device='cpu',
                     dtype=torch.float32, sparse=False):
    """Camera + extra-modality tensors ~ N(0,1).  `sparse=True` makes the extra
    modalities mostly a constant background like projected lidar/radar images."""
    g = torch.Generator().manual_seed(1000 + seed)
    cam = torch.randn(batch, 3, H, W, generator=g)
    mods = []
    for c in mod_channels:
        m = torch.randn(batch, c, H, W, generator=g)
        if sparse:
            keep = torch.rand(batch, 1, H, W, generator=g) < 0.05
            m = torch.where(keep, m, torch.full_like(m, -0.5))
        mods.append(m)
    return cam.to(device=device, dtype=dtype), [m.to(device=device, dtype=dtype) for m in mods]


def rel_err(a, b):
    """norm-wise relative error ||a-b|| / ||b||"""
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
