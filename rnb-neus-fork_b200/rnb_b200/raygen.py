"""Device-resident ray batches.

`DeviceRayBatcher` holds what `models/dataset.py` keeps on the CPU -- images, warm-up images, masks, per-pixel light
directions, K^-1 and poses -- in HBM, and gathers one training batch with a single kernel
(reference Dataset.ps_gen_random_rays_at_view_on_all_lights, models/dataset.py:351-376; near_far_from_sphere :448-458;
light gather exp_runner.py:214-220).  Pixel indices come from `torch.randint(..., device='cpu')` with the same call order
as the reference, so a seeded run draws the same pixels.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as L


class DeviceRayBatcher:
    def __init__(self, images, images_warmup, masks, light_directions, intrinsics_all_inv, pose_all, device="cuda"):
        """images / images_warmup [V,L,H,W,3], masks [V,H,W,C], light_directions [V,L,H,W,3] (or None),
        intrinsics_all_inv / pose_all [V,4,4]"""
        f = lambda t: None if t is None else t.detach().to(device=device, dtype=torch.float32).contiguous()
        self.images, self.images_warmup, self.masks = f(images), f(images_warmup), f(masks)
        self.light_directions = f(light_directions)
        self.intrinsics_all_inv, self.pose_all = f(intrinsics_all_inv), f(pose_all)
        L.require_cuda(self.images, "DeviceRayBatcher")
        self.n_images, self.n_lights, self.H, self.W = self.images.shape[:4]
        self.device = self.images.device
        self._pix_ring, self._pix_slot = [], 0

    def _pixels_to_device(self, pixels_x, pixels_y):
        """CPU pixel indices (the reference draws them with the CPU generator) -> device, without blocking the host: a
        pageable-memory copy would make every iteration wait for the GPU to drain; a small ring of pinned staging
        buffers, each guarded by the event of its last copy, keeps the launch queue running ahead."""
        B = pixels_x.numel()
        if not self._pix_ring or self._pix_ring[0][0].shape[1] != B:
            self._pix_ring = [(torch.empty(2, B, dtype=torch.int64).pin_memory(), torch.cuda.Event()) for _ in range(4)]
            self._pix_slot = 0
            for _, ev in self._pix_ring:
                ev.record()
        host, ev = self._pix_ring[self._pix_slot]
        self._pix_slot = (self._pix_slot + 1) % len(self._pix_ring)
        ev.synchronize()
        host[0].copy_(pixels_x)
        host[1].copy_(pixels_y)
        d = host.to(self.device, non_blocking=True)
        ev.record()
        return d[0], d[1]

    def gather(self, img_idx, pixels_x, pixels_y, want_lights=True):
        """-> dict(rays_o, rays_d [B,3], near, far, mask [B,1], images_warmup, images [L,B,3], lights_dir [L,B,1,3])"""
        dev = self.device
        if not pixels_x.is_cuda and not pixels_y.is_cuda and pixels_x.dim() == 1 and pixels_x.shape == pixels_y.shape:
            px, py = self._pixels_to_device(pixels_x, pixels_y)
        else:
            px = pixels_x.to(dev, torch.int64).contiguous()
            py = pixels_y.to(dev, torch.int64).contiguous()
        B = px.numel()
        f32 = dict(dtype=torch.float32, device=dev)
        out = dict(rays_o=torch.empty(B, 3, **f32), rays_d=torch.empty(B, 3, **f32), near=torch.empty(B, 1, **f32),
                   far=torch.empty(B, 1, **f32), mask=torch.empty(B, 1, **f32),
                   images_warmup=torch.empty(self.n_lights, B, 3, **f32), images=torch.empty(self.n_lights, B, 3, **f32))
        lights = None
        if want_lights and self.light_directions is not None:
            lights = torch.empty(self.n_lights, B, 3, **f32)
        p = L.RayBatch()
        p.n_rays, p.n_lights, p.H, p.W = B, self.n_lights, self.H, self.W
        p.intrinsics_inv = L.ptr(self.intrinsics_all_inv[img_idx])
        p.pose = L.ptr(self.pose_all[img_idx])
        p.pixels_x, p.pixels_y = L.ptr(px), L.ptr(py)
        p.images, p.images2 = L.ptr(self.images_warmup[img_idx]), L.ptr(self.images[img_idx])
        p.mask, p.mask_channels = L.ptr(self.masks[img_idx]), self.masks.shape[-1]
        p.light_dirs = L.ptr(self.light_directions[img_idx]) if lights is not None else None
        p.rays_o, p.rays_d, p.near, p.far = (L.ptr(out[k]) for k in ("rays_o", "rays_d", "near", "far"))
        p.mask_out, p.rgb, p.rgb2, p.lights = L.ptr(out["mask"]), L.ptr(out["images_warmup"]), L.ptr(out["images"]), L.ptr(lights)
        L.check(L.load().rnb_ray_batch(C.byref(p), L.stream_ptr()), "ray_batch")
        out["lights_dir"] = lights.view(self.n_lights, B, 1, 3) if lights is not None else None
        return out

    def ps_gen_random_rays_at_view_on_all_lights(self, img_idx, batch_size):
        """Same return tuple as the reference method: (cat[rays_o, rays_v, mask] [B,7], images_warmup [L,B,3],
        images [L,B,3], pixels_x, pixels_y); the RNG calls (CPU randint for x, then y) match models/dataset.py:356-357."""
        pixels_x = torch.randint(low=0, high=self.W, size=[batch_size], device='cpu')
        pixels_y = torch.randint(low=0, high=self.H, size=[batch_size], device='cpu')
        g = self.gather(img_idx, pixels_x, pixels_y, want_lights=False)
        data = torch.cat([g["rays_o"], g["rays_d"], g["mask"]], dim=-1)
        return data, g["images_warmup"], g["images"], pixels_x.to(self.device), pixels_y.to(self.device)

    def near_far_from_sphere(self, rays_o, rays_d):
        """reference models/dataset.py:448-458 (tiny elementwise expression; kept in torch for callers that have rays only)"""
        a = torch.sum(rays_d ** 2, dim=-1, keepdim=True)
        b = 2.0 * torch.sum(rays_o * rays_d, dim=-1, keepdim=True)
        mid = 0.5 * (-b) / a
        return mid - 1.0, mid + 1.0
