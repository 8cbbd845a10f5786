"""Thin functional wrappers: torch tensors in, one C-ABI call each (no autograd here; see ops.py)."""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as L

N_SDF_LAYERS = 9


def _points(n_pts, x=None, rays_o=None, rays_d=None, z=None, n_per_ray=0, grid_res=0, slab_x0=0,
            bmin=(0, 0, 0), bmax=(0, 0, 0)):
    p = L.Points()
    p.n_pts = int(n_pts)
    p.x = L.ptr(x)
    p.rays_o = L.ptr(rays_o)
    p.rays_d = L.ptr(rays_d)
    p.z = L.ptr(z)
    p.n_per_ray = int(n_per_ray)
    p.grid_res = int(grid_res)
    p.slab_x0 = int(slab_x0)
    for a in range(3):
        p.bmin[a] = float(bmin[a])
        p.bmax[a] = float(bmax[a])
    p._keep = (x, rays_o, rays_d, z)
    return p


def points_explicit(x):
    x = x.detach().float().contiguous()
    L.require_cuda(x, "points")
    return _points(x.shape[0], x=x)


def points_rays(rays_o, rays_d, z):
    """z [B, n]: point (b, i) = o_b + d_b * z[b, i]"""
    o = rays_o.detach().float().contiguous()
    d = rays_d.detach().float().contiguous()
    z = z.detach().float().contiguous()
    L.require_cuda(z, "points")
    return _points(z.numel(), rays_o=o, rays_d=d, z=z, n_per_ray=z.shape[1])


def points_grid(bound_min, bound_max, resolution, x0, nx):
    return _points(int(nx) * resolution * resolution, grid_res=resolution, slab_x0=x0,
                   bmin=[float(v) for v in bound_min], bmax=[float(v) for v in bound_max])


class SdfPacked:
    """Packed fp16 operand images + fp32 side table of one SDFNetwork state."""

    def __init__(self, device):
        lib = L.load()
        self.wblob = torch.empty(lib.rnb_sdf_wblob_bytes(), dtype=torch.uint8, device=device)
        self.aux = torch.empty(lib.rnb_sdf_aux_floats(), dtype=torch.float32, device=device)

    def pack(self, Ws, bs):
        assert len(Ws) == N_SDF_LAYERS and len(bs) == N_SDF_LAYERS
        expect = [(256, 39), (256, 256), (256, 256), (217, 256), (256, 256), (256, 256), (256, 256), (256, 256),
                  (257, 256)]
        keep = []
        wp = (C.c_void_p * N_SDF_LAYERS)()
        bp = (C.c_void_p * N_SDF_LAYERS)()
        for l in range(N_SDF_LAYERS):
            W = Ws[l].detach().float().contiguous()
            b = bs[l].detach().float().contiguous()
            if tuple(W.shape) != expect[l]:
                raise RuntimeError(f"rnb_b200: SDF layer {l} has shape {tuple(W.shape)}, the sm_100a kernels are "
                                   f"specialised for the shipped 8x256 network ({expect[l]})")
            keep += [W, b]
            wp[l] = L.ptr(W)
            bp[l] = L.ptr(b)
        L.check(L.load().rnb_sdf_pack(wp, bp, L.ptr(self.wblob), L.ptr(self.aux), L.stream_ptr()), "sdf_pack")
        self._keep = keep
        return self


def sdf_fwd(packed: SdfPacked, pts, out=None, out_scale=1.0):
    if out is None:
        out = torch.empty(pts.n_pts, dtype=torch.float32, device=packed.wblob.device)
    L.check(L.load().rnb_sdf_fwd(C.byref(pts), L.ptr(packed.wblob), L.ptr(packed.aux), L.ptr(out), float(out_scale),
                                 L.stream_ptr()), "sdf_fwd")
    return out


class SdfStreams:
    """fp16 activation streams written by sdf_fwd_grad and consumed by the backward kernels."""

    def __init__(self, n_pts, device):
        lib = L.load()
        self.n_pts = n_pts
        self.n_pad = lib.rnb_padded_points(n_pts)
        s256 = lib.rnb_stream_bytes(n_pts, 256)
        s64 = lib.rnb_stream_bytes(n_pts, 64)
        u8 = dict(dtype=torch.uint8, device=device)
        self.stride = s256
        self.feat = torch.empty(s256, **u8)
        self.in0 = torch.empty(s64, **u8)
        self.inl = torch.empty(8 * s256, **u8)
        self.s = torch.empty(8 * s256, **u8)
        self.w = torch.empty(8 * s256, **u8)


def sdf_fwd_grad(packed: SdfPacked, pts, streams: SdfStreams = None, want_full=False):
    dev = packed.wblob.device
    n = pts.n_pts
    if streams is None:
        streams = SdfStreams(n, dev)
    sdf = torch.empty(n, dtype=torch.float32, device=dev)
    grad = torch.empty(n, 3, dtype=torch.float32, device=dev)
    full = torch.empty(n, 257, dtype=torch.float32, device=dev) if want_full else None
    L.check(L.load().rnb_sdf_fwd_grad(C.byref(pts), L.ptr(packed.wblob), L.ptr(packed.aux), L.ptr(sdf), L.ptr(grad),
                                      L.ptr(full), L.ptr(streams.feat), L.ptr(streams.in0), L.ptr(streams.inl),
                                      L.ptr(streams.s), L.ptr(streams.w), L.stream_ptr()), "sdf_fwd_grad")
    return sdf, grad, full, streams


SDF_W_SHAPES = [(256, 39), (256, 256), (256, 256), (217, 256), (256, 256), (256, 256), (256, 256), (256, 256), (257, 256)]


def sdf_bwd(packed: SdfPacked, pts, streams: SdfStreams, d_sdf, d_grad, d_feat=None, scratch=None):
    """-> (dWs, dbs): gradients w.r.t. the effective (weight-norm folded) weights and biases."""
    dev = packed.wblob.device
    lib = L.load()
    n = pts.n_pts
    need = lib.rnb_sdf_bwd_scratch_bytes(n)
    if scratch is None or scratch.numel() < need:
        scratch = torch.empty(need, dtype=torch.uint8, device=dev)
    d_sdf = d_sdf.detach().float().contiguous().view(-1)
    d_grad = d_grad.detach().float().contiguous().view(-1, 3)
    if d_feat is not None:
        d_feat = d_feat.detach().float().contiguous().view(-1, 256)
    dWs = [torch.empty(s, dtype=torch.float32, device=dev) for s in SDF_W_SHAPES]
    dbs = [torch.empty(s[0], dtype=torch.float32, device=dev) for s in SDF_W_SHAPES]
    wp = (C.c_void_p * N_SDF_LAYERS)(*[L.ptr(t) for t in dWs])
    bp = (C.c_void_p * N_SDF_LAYERS)(*[L.ptr(t) for t in dbs])
    L.check(lib.rnb_sdf_bwd(C.byref(pts), L.ptr(packed.wblob), L.ptr(packed.aux), L.ptr(d_sdf), L.ptr(d_grad),
                            L.ptr(d_feat), L.ptr(streams.in0), L.ptr(streams.inl), L.ptr(streams.s), L.ptr(streams.w),
                            L.ptr(scratch), wp, bp, L.stream_ptr()), "sdf_bwd")
    return dWs, dbs, scratch


def stream_to_rowmajor(buf, n_pts, cols, dtype=torch.float16):
    """Decode a stream image [Npad/64][cols/8][64][8] into [n_pts, cols] (tests / debugging)."""
    t = buf.view(dtype).view(-1, cols // 8, 64, 8).permute(0, 2, 1, 3).reshape(-1, cols)
    return t[:n_pts]
