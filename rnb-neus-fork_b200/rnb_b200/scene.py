"""Analytic DiLiGenT-MV-shaped scene for end-to-end runs without data files.

Produces the tensors `models/dataset.py` builds from PNG normal / albedo / mask maps and `cameras.npz`
(reference models/dataset.py:130-240), for a sphere of known radius seen by a ring of cameras:

  images_warmup [V,L,H,W,3] = albedo * max(n . l_warmup, 0),   l_warmup: one direction per view (tilt 0/120/240 deg,
                              slant 30 deg around the viewing axis, gen_light_directions() :253-265)
  images        [V,L,H,W,3] = albedo * max(n . l_pixel, 0),    l_pixel: the same three lights re-centred on every
                              pixel's normal with slant 54.74 deg (gen_light_directions(normals) :267-289, restated
                              with an explicit tangent frame instead of the per-pixel SVD)
  masks [V,H,W,1], light_directions [V,L,H,W,3], light_directions_warmup [V,L,3], intrinsics_all_inv / pose_all [V,4,4]

so that `DeviceRayBatcher` + `NeuSRenderer.render_rnb[_warmup]` + the loss of exp_runner.py:241-256 can be driven exactly
like `Runner.train_rnb` does.  Host-side setup code (plain torch); nothing here is on the hot path.
"""
from __future__ import annotations

import math

import torch


def _look_at(cam, target=(0.0, 0.0, 0.0)):
    """camera-to-world pose [4,4]; camera looks along +z, y down (OpenCV convention, like load_K_Rt_from_P)."""
    c = torch.tensor(cam, dtype=torch.float64)
    z = torch.tensor(target, dtype=torch.float64) - c
    z = z / z.norm()
    up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64)
    x = torch.linalg.cross(z, up)
    x = x / x.norm()
    y = torch.linalg.cross(z, x)
    pose = torch.eye(4, dtype=torch.float64)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = x, y, z, c
    return pose


def default_albedo(p):
    """smooth position-dependent albedo in [0.35, 0.95]"""
    return 0.65 + 0.3 * torch.stack([torch.sin(3.0 * p[..., 0]), torch.sin(3.0 * p[..., 1] + 1.0),
                                     torch.sin(3.0 * p[..., 2] + 2.0)], dim=-1)


def sphere_scene(n_views=8, H=64, W=64, radius=0.6, cam_dist=3.0, elevation=0.35, albedo_fn=default_albedo):
    """-> dict of CPU float32 tensors with the shapes of the reference Dataset attributes of the same names."""
    tilt = torch.tensor([0.0, 120.0, 240.0], dtype=torch.float64) * math.pi / 180.0
    n_lights = tilt.numel()

    def lights_local(slant_deg):          # [L,3] around +z of a local frame, pointing away from the surface
        s = math.radians(slant_deg)
        return torch.stack([math.sin(s) * torch.cos(tilt), math.sin(s) * torch.sin(tilt),
                            torch.full_like(tilt, math.cos(s))], dim=-1)

    # the sphere's silhouette (angular radius asin(r / D)) fills ~70 % of the half field of view
    focal = 0.5 * W / math.tan(math.asin(radius / cam_dist) / 0.7)
    K = torch.eye(4, dtype=torch.float64)
    K[0, 0] = K[1, 1] = focal
    K[0, 2], K[1, 2] = 0.5 * (W - 1), 0.5 * (H - 1)
    Kinv = torch.linalg.inv(K)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    pix = torch.stack([xs, ys, torch.ones_like(xs)], dim=-1)                     # [H,W,3]
    dirs_cam = pix @ Kinv[:3, :3].T
    dirs_cam = dirs_cam / dirs_cam.norm(dim=-1, keepdim=True)

    out = dict(images=[], images_warmup=[], masks=[], light_directions=[], light_directions_warmup=[], pose_all=[])
    lw_cam = lights_local(30.0) * torch.tensor([1.0, 1.0, -1.0], dtype=torch.float64)   # towards the camera (-z)
    lp_loc = lights_local(54.74)
    for v in range(n_views):
        phi = 2.0 * math.pi * v / n_views
        cam = (cam_dist * math.cos(phi) * math.cos(elevation), cam_dist * math.sin(phi) * math.cos(elevation),
               cam_dist * math.sin(elevation))
        pose = _look_at(cam)
        R, o = pose[:3, :3], pose[:3, 3]
        d = dirs_cam @ R.T                                                     # [H,W,3] world
        b = (d * o).sum(-1)
        disc = b * b - (o @ o - radius * radius)
        hit = disc > 0
        t = -b - torch.sqrt(disc.clamp_min(0.0))
        p = o + d * t[..., None]
        n = torch.where(hit[..., None], p / radius, torch.zeros_like(p))       # world normals, 0 off the object
        alb = torch.where(hit[..., None], albedo_fn(p), torch.zeros_like(p))
        lw = lw_cam @ R.T                                                      # [L,3] world
        # per-pixel lights: tangent frame (t1, t2, n)
        a = torch.where((n[..., 2].abs() < 0.9)[..., None], torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64),
                        torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64))
        t1 = torch.linalg.cross(a.expand_as(n), n)
        t1 = t1 / t1.norm(dim=-1, keepdim=True).clamp_min(1e-12)
        t2 = torch.linalg.cross(n, t1)
        lp = (lp_loc[:, None, None, 0:1] * t1 + lp_loc[:, None, None, 1:2] * t2 + lp_loc[:, None, None, 2:3] * n)
        lp = torch.where(hit[None, ..., None], lp, lw[:, None, None, :].expand_as(lp))       # [L,H,W,3]
        sh_w = (n[None] * lw[:, None, None, :]).sum(-1).clamp_min(0.0)[..., None]            # [L,H,W,1]
        sh_p = (n[None] * lp).sum(-1).clamp_min(0.0)[..., None]
        out["images_warmup"].append(alb[None] * sh_w)
        out["images"].append(alb[None] * sh_p)
        out["masks"].append(hit[..., None].to(torch.float64))
        out["light_directions"].append(lp)
        out["light_directions_warmup"].append(lw)
        out["pose_all"].append(pose)
    res = {k: torch.stack(vs).float() for k, vs in out.items()}
    res["intrinsics_all_inv"] = Kinv.float()[None].repeat(n_views, 1, 1)
    res.update(n_images=n_views, n_lights=n_lights, H=H, W=W, radius=radius)
    return res


def learning_rate_factor(iter_step, warm_up_end, end_iter, alpha):
    """Runner.update_learning_rate (exp_runner.py:320-332): linear warm-up, then cosine decay to alpha."""
    if iter_step < warm_up_end:
        return iter_step / warm_up_end
    progress = (iter_step - warm_up_end) / (end_iter - warm_up_end)
    return (math.cos(math.pi * progress) + 1.0) * 0.5 * (1 - alpha) + alpha
