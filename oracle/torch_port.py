"""TEST / MEASUREMENT INFRASTRUCTURE -- never imported by the product path.

A second CPU restatement of the reference's train_rnb hot path, this time the way the reference itself computes it:
float32 torch tensors and *autograd* (the eikonal gradient through `autograd.grad(..., create_graph=True)`, the
parameter gradients through `loss.backward()`), so that it costs what the reference costs on host cores.  The numpy
float64 oracle (rnb_oracle.py) stays the parity checker; this port is the CPU arm of `bench.py` on machines where the
reference tree is absent (the GPU box), where the numpy port would understate the reference by ~10x.

Pinned against the reference's own outputs (tests/golden/render_*.npz) in tests/test_oracle_golden.py.
Every function cites the reference lines it restates.  State dicts use the reference's key names
(`lin{l}.weight_g / weight_v / bias`, `variance`).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def embed(x, multires):
    """models/embedder.py:21-55: [x, sin(2^k x), cos(2^k x)]_k with freq = 2**linspace(0, L-1, L)"""
    out = [x]
    for k in range(multires):
        f = 2.0 ** k
        out += [torch.sin(x * f), torch.cos(x * f)]
    return torch.cat(out, -1)


def fold(sd, l):
    """weight-norm: W = g * v / |v|_row   (models/fields.py:72-74, torch.nn.utils.weight_norm dim=0)"""
    v, g = sd[f"lin{l}.weight_v"], sd[f"lin{l}.weight_g"]
    return g * v / v.norm(dim=1, keepdim=True)


def sdf_forward(sd, x, multires=6, skip=4, n_lin=9):
    """SDFNetwork.forward, models/fields.py:82-104 (scale = 1)"""
    e = embed(x, multires)
    h = e
    for l in range(n_lin):
        if l == skip:
            h = torch.cat([h, e], -1) / math.sqrt(2.0)
        h = F.linear(h, fold(sd, l), sd[f"lin{l}.bias"])
        if l < n_lin - 1:
            h = F.softplus(h, beta=100)
    return h


def sdf_gradient(sd, x):
    """SDFNetwork.gradient, models/fields.py:114-127: a second forward + autograd.grad(create_graph=True)"""
    x = x.detach().requires_grad_(True)
    y = sdf_forward(sd, x)[:, :1]
    return torch.autograd.grad(y, x, torch.ones_like(y), create_graph=True, retain_graph=True, only_inputs=True)[0]


def color_forward(sd, points, normals, feat, multires_view=4):
    """RenderingNetwork.forward mode 'no_view_dir', models/fields.py:177-215"""
    h = torch.cat([embed(points, multires_view), embed(normals, multires_view), feat], -1)
    for l in range(3):
        h = F.linear(h, fold(sd, l), sd[f"lin{l}.bias"])
        if l < 2:
            h = F.relu(h)
    return torch.sigmoid(h)


def sample_pdf(bins, weights, n_samples):
    """models/renderer.py:39-69, det=True"""
    weights = weights + 1e-5
    pdf = weights / weights.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
    u = torch.linspace(0.5 / n_samples, 1.0 - 0.5 / n_samples, n_samples).expand(cdf.shape[0], n_samples).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below, above = (inds - 1).clamp(min=0), inds.clamp(max=cdf.shape[-1] - 1)
    cb, ca = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bb, ba = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = ca - cb
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    return bb + (u - cb) / denom * (ba - bb)


def up_sample(rays_o, rays_d, z, sdf, n_importance, inv_s):
    """models/renderer.py:132-176"""
    B = z.shape[0]
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z[..., :, None]
    radius = pts.norm(dim=-1)
    inside = (radius[:, :-1] < 1.0) | (radius[:, 1:] < 1.0)
    ps, ns, pz, nz = sdf[:, :-1], sdf[:, 1:], z[:, :-1], z[:, 1:]
    mid = (ps + ns) * 0.5
    cos = (ns - ps) / (nz - pz + 1e-5)
    prev = torch.cat([torch.zeros(B, 1), cos[:, :-1]], -1)
    cos = torch.minimum(prev, cos).clip(-1e3, 0.0) * inside
    dist = nz - pz
    pc = torch.sigmoid((mid - cos * dist * 0.5) * inv_s)
    nc = torch.sigmoid((mid + cos * dist * 0.5) * inv_s)
    alpha = (pc - nc + 1e-5) / (pc + 1e-5)
    w = alpha * torch.cumprod(torch.cat([torch.ones(B, 1), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    return sample_pdf(z, w, n_importance).detach()


def render_rnb(sdf_sd, col_sd, variance, rays_o, rays_d, near, far, lights, t_rand, cos_anneal_ratio=1.0, warmup=True,
               no_albedo=False, n_samples=64, n_importance=64, up_sample_steps=4, z_vals=None):
    """render_rnb_warmup / render_rnb, models/renderer.py:828-930 / 932-1033, with render_core_mvps (:466-554)"""
    B = rays_o.shape[0]
    sample_dist = 2.0 / n_samples
    if z_vals is None:
        with torch.no_grad():
            z = near + (far - near) * torch.linspace(0.0, 1.0, n_samples)[None, :]
            if t_rand is not None:
                z = z + t_rand * 2.0 / n_samples
            sdf = sdf_forward(sdf_sd, (rays_o[:, None, :] + rays_d[:, None, :] * z[..., :, None]).reshape(-1, 3))[:, :1].reshape(B, -1)
            for i in range(up_sample_steps):
                new_z = up_sample(rays_o, rays_d, z, sdf, n_importance // up_sample_steps, 64 * 2 ** i)
                zc = torch.cat([z, new_z], -1)
                zc, idx = torch.sort(zc, -1)
                if i + 1 < up_sample_steps:
                    new_sdf = sdf_forward(sdf_sd, (rays_o[:, None, :] + rays_d[:, None, :] * new_z[..., :, None]).reshape(-1, 3))[:, :1]
                    sdf = torch.gather(torch.cat([sdf, new_sdf.reshape(B, -1)], -1), 1, idx)
                z = zc
    else:
        z = z_vals
    n = z.shape[1]
    dists = torch.cat([z[..., 1:] - z[..., :-1], torch.full((B, 1), sample_dist)], -1)
    mid = z + dists * 0.5
    pts = (rays_o[:, None, :] + rays_d[:, None, :] * mid[..., :, None]).reshape(-1, 3)
    dirs = rays_d[:, None, :].expand(B, n, 3).reshape(-1, 3)
    out = sdf_forward(sdf_sd, pts)
    sdf, feat = out[:, :1], out[:, 1:]
    grad = sdf_gradient(sdf_sd, pts)
    albedo = torch.ones(B * n, 3) if no_albedo else color_forward(col_sd, pts, grad, feat)
    inv_s = torch.exp(variance * 10.0).clip(1e-6, 1e6).expand(B * n, 1)
    true_cos = (dirs * grad).sum(-1, keepdim=True)
    r = cos_anneal_ratio
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - r) + F.relu(-true_cos) * r)
    d1 = dists.reshape(-1, 1)
    pc = torch.sigmoid((sdf - iter_cos * d1 * 0.5) * inv_s)
    nc = torch.sigmoid((sdf + iter_cos * d1 * 0.5) * inv_s)
    alpha = ((pc - nc + 1e-5) / (pc + 1e-5)).reshape(B, n).clip(0.0, 1.0)
    pn = pts.norm(dim=-1).reshape(B, n)
    inside = (pn < 1.0).float()
    relax = (pn < 1.2).float()
    w = alpha * torch.cumprod(torch.cat([torch.ones(B, 1), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    g3 = grad.reshape(B, n, 3)
    eik = (relax * (g3.norm(dim=-1) - 1.0) ** 2).sum() / (relax.sum() + 1e-5)
    # RNb shading (renderer.py:904-918 / 1008-1017)
    L = lights.shape[0]
    shade = (g3[None] * lights.expand(L, B, 1, 3)).sum(-1)
    if warmup:
        shade = F.relu(shade)
    color = (albedo.reshape(1, B, n, 3) * w[None, :, :, None] * shade[..., None]).sum(2)
    return dict(color_fine=color, weight_sum=w.sum(-1, keepdim=True), weights=w, gradient_error=eik, gradients=g3,
                cdf_fine=pc.reshape(B, n), inside_sphere=inside, weight_max=w.max(-1, keepdim=True)[0],
                s_val=(1.0 / inv_s).reshape(B, n).mean(-1, keepdim=True), z_vals=z)


def loss_fn(out, true_rgb, mask, igr_weight=0.1, mask_weight=0.1):
    """exp_runner.py:241-256"""
    mask_sum = mask.sum() + 1e-5
    err = ((out["color_fine"] - true_rgb) * mask[None, :, :]).reshape(-1, 3)
    color = F.l1_loss(err, torch.zeros_like(err), reduction="sum") / (mask_sum * true_rgb.shape[0])
    return color + out["gradient_error"] * igr_weight + F.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), mask) * mask_weight


def extract_block(sdf_sd, bmin, bmax, R):
    """extract_fields, models/renderer.py:10-25 (one block)"""
    X = [torch.linspace(float(bmin[a]), float(bmax[a]), R) for a in range(3)]
    xx, yy, zz = torch.meshgrid(X[0], X[1], X[2], indexing="ij")
    pts = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], -1)
    with torch.no_grad():
        return -sdf_forward(sdf_sd, pts)[:, 0].reshape(R, R, R)
