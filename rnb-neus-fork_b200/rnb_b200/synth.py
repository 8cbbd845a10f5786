"""Synthetic DiLiGenT-MV / bearPNG-shaped ray batches.

One view per step: all rays share one camera origin at ||c|| ~ 3, unit directions
toward uniform targets in a +-0.6 cube (every ray hits the unit sphere),
near/far = -(o.d) -+ 1 (reference models/dataset.py:448-458), L = 3 unit light
directions ([L,1,1,3] in warm-up, [L,B,1,3] afterwards, exp_runner.py:203-220),
true_rgb ~ U[0,1), mask ~ Bernoulli(0.7).  Everything is drawn from a CPU
generator so the same seed gives the same batch on every machine.
"""
from __future__ import annotations

import torch

WMASK_CONF = dict(
    # confs/wmask_rnb.conf:41-90 (identical model block in all four shipped confs)
    nerf=dict(D=8, d_in=4, d_in_view=3, W=256, multires=10, multires_view=4, output_ch=4, skips=[4],
              use_viewdirs=True),
    sdf_network=dict(d_out=257, d_in=3, d_hidden=256, n_layers=8, skip_in=[4], multires=6, bias=0.5,
                     scale=1.0, geometric_init=True, weight_norm=True),
    variance_network=dict(init_val=0.3),
    rendering_network=dict(d_feature=256, mode="no_view_dir", d_in=6, d_out=3, d_hidden=256, n_layers=2,
                           weight_norm=True, multires_view=4, squeeze_out=True),
    neus_renderer=dict(n_samples=64, n_importance=64, n_outside=0, up_sample_steps=4, perturb=1.0),
    igr_weight=0.1, mask_weight=0.1,
)


# 'trained-like' perturbation used by the golden fixtures, tests and bench
SDF_NOISE = 0.004
COLOR_NOISE = 0.03
TRAINED_VARIANCE = 0.45     # inv_s = exp(4.5) ~ 90


def make_batch(batch_size: int, n_lights: int = 3, warmup: bool = True, seed: int = 1, view: int = 0):
    """Returns a dict of CPU float32 tensors."""
    g = torch.Generator().manual_seed(seed * 1000003 + view)
    c = torch.randn(3, generator=g)
    c = 3.0 * c / c.norm()
    target = (torch.rand(batch_size, 3, generator=g) - 0.5) * 1.2
    rays_o = c.expand(batch_size, 3).contiguous()
    rays_d = target - rays_o
    rays_d = rays_d / rays_d.norm(dim=-1, keepdim=True)
    a = (rays_d ** 2).sum(-1, keepdim=True)
    b = 2.0 * (rays_o * rays_d).sum(-1, keepdim=True)
    mid = 0.5 * (-b) / a
    near, far = mid - 1.0, mid + 1.0
    if warmup:
        lights = torch.randn(n_lights, 1, 1, 3, generator=g)
    else:
        lights = torch.randn(n_lights, batch_size, 1, 3, generator=g)
    lights = lights / lights.norm(dim=-1, keepdim=True)
    true_rgb = torch.rand(n_lights, batch_size, 3, generator=g)
    mask = (torch.rand(batch_size, 1, generator=g) < 0.7).float()
    t_rand = torch.rand(batch_size, 1, generator=g) - 0.5
    return dict(rays_o=rays_o, rays_d=rays_d, near=near, far=far, lights_dir=lights, true_rgb=true_rgb,
                mask=mask, t_rand=t_rand)


def perturb_state_dict_(module: torch.nn.Module, std: float = 0.03, seed: int = 5):
    """'Trained-like' weights: geometric init + N(0, std) noise on every weight_v / bias."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(module.named_parameters()):
            if name.endswith("weight_v") or name.endswith(".weight"):
                p.add_(torch.randn(p.shape, generator=g).to(p.device) * std)
            elif name.endswith("bias"):
                p.add_(torch.randn(p.shape, generator=g).to(p.device) * (std * 0.3))
    return module
