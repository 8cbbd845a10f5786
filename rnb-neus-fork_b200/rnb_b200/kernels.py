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

    def __init__(self, n_pts, device, for_backward=True):
        lib = L.load()
        self.n_pts = n_pts
        self.n_pad = lib.rnb_padded_points(n_pts)
        s256 = lib.rnb_stream_bytes(n_pts, 256)
        s64 = lib.rnb_stream_bytes(n_pts, 64)
        u8 = dict(dtype=torch.uint8, device=device)
        self.stride = s256
        self.feat = torch.empty(s256, **u8)
        # in0 and w are read only by the backward: an inference pass neither allocates nor writes them
        self.in0 = torch.empty(s64, **u8) if for_backward else None
        self.inl = torch.empty(8 * s256, **u8)
        self.w = torch.empty(8 * s256, **u8) if for_backward else None


def sdf_fwd_grad(packed: SdfPacked, pts, streams: SdfStreams = None, want_full=False, for_backward=True):
    dev = packed.wblob.device
    n = pts.n_pts
    if streams is None:
        streams = SdfStreams(n, dev, for_backward)
    sdf = torch.empty(n, dtype=torch.float32, device=dev)
    grad = torch.empty(n, 3, dtype=torch.float32, device=dev)
    full = torch.empty(n, 257, dtype=torch.float32, device=dev) if want_full else None
    L.check(L.load().rnb_sdf_fwd_grad(C.byref(pts), L.ptr(packed.wblob), L.ptr(packed.aux), L.ptr(sdf), L.ptr(grad),
                                      L.ptr(full), L.ptr(streams.feat), L.ptr(streams.in0), L.ptr(streams.inl),
                                      L.ptr(streams.w), L.stream_ptr()), "sdf_fwd_grad")
    return sdf, grad, full, streams


SDF_W_SHAPES = [(256, 39), (256, 256), (256, 256), (217, 256), (256, 256), (256, 256), (256, 256), (256, 256), (257, 256)]


def sdf_bwd(packed: SdfPacked, pts, streams: SdfStreams, d_sdf, d_grad, d_feat=None, scratch=None, d_feat16=None):
    """-> (dWs, dbs): gradients w.r.t. the effective (weight-norm folded) weights and biases.
    d_feat: fp32 [n,256], or d_feat16 = (fp16 stream, meta float[2]) as written by the albedo backward."""
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
    f16, meta = d_feat16 if d_feat16 is not None else (None, None)
    L.check(lib.rnb_sdf_bwd(C.byref(pts), L.ptr(packed.wblob), L.ptr(packed.aux), L.ptr(d_sdf), L.ptr(d_grad),
                            L.ptr(d_feat), L.ptr(f16), L.ptr(meta), L.ptr(streams.in0), L.ptr(streams.inl), L.ptr(streams.w),
                            L.ptr(scratch), wp, bp, L.stream_ptr()), "sdf_bwd")
    return dWs, dbs, scratch


# ----------------------------------------------------------------------------- per-ray kernels
def _f32(t):
    return t.detach().float().contiguous()


def coarse_z(near, far, t_rand, n_samples):
    near, far = _f32(near).view(-1), _f32(far).view(-1)
    L.require_cuda(near, "coarse_z")
    tr = _f32(t_rand).view(-1) if t_rand is not None else None
    z = torch.empty(near.numel(), n_samples, dtype=torch.float32, device=near.device)
    L.check(L.load().rnb_coarse_z(L.ptr(near), L.ptr(far), L.ptr(tr), L.ptr(z), near.numel(), n_samples, L.stream_ptr()),
            "coarse_z")
    return z


def upsample_step(rays_o, rays_d, z_old, sdf_old, inv_s, n_new, z_pending=None, sdf_pending=None, want_merged=True,
                  want_debug=False):
    """-> (z_new [B,n_new], z_merged, sdf_merged, inds, cdf)"""
    B, n_old = z_old.shape
    n_merge = 0 if z_pending is None else z_pending.shape[1]
    dev = z_old.device
    f32 = dict(dtype=torch.float32, device=dev)
    z_new = torch.empty(B, n_new, **f32)
    z_m = torch.empty(B, n_old + n_merge, **f32) if (want_merged and n_merge) else None
    s_m = torch.empty(B, n_old + n_merge, **f32) if (want_merged and n_merge) else None
    inds = torch.empty(B, n_new, dtype=torch.int32, device=dev) if want_debug else None
    cdf = torch.empty(B, n_old + n_merge, **f32) if want_debug else None
    p = L.Upsample()
    p.n_rays = B
    p.rays_o, p.rays_d = L.ptr(rays_o), L.ptr(rays_d)
    p.z_old, p.sdf_old, p.n_old = L.ptr(z_old), L.ptr(sdf_old), n_old
    p.z_pending, p.sdf_pending, p.n_merge = L.ptr(z_pending), L.ptr(sdf_pending), n_merge
    p.z_merged, p.sdf_merged = L.ptr(z_m), L.ptr(s_m)
    p.inv_s, p.n_new = float(inv_s), n_new
    p.z_new, p.inds, p.cdf_out = L.ptr(z_new), L.ptr(inds), L.ptr(cdf)
    L.check(L.load().rnb_upsample_step(C.byref(p), L.stream_ptr()), "upsample_step")
    if not n_merge:
        z_m, s_m = z_old, sdf_old
    return z_new, z_m, s_m, inds, cdf


def sample_pdf_from_cdf(bins, cdf, n_new):
    bins, cdf = _f32(bins), _f32(cdf)
    B, n = bins.shape
    samples = torch.empty(B, n_new, dtype=torch.float32, device=bins.device)
    inds = torch.empty(B, n_new, dtype=torch.int64, device=bins.device)
    L.check(L.load().rnb_sample_pdf_from_cdf(L.ptr(bins), L.ptr(cdf), B, n, n_new, L.ptr(samples), L.ptr(inds),
                                             L.stream_ptr()), "sample_pdf_from_cdf")
    return samples, inds


def final_merge(z_old, z_new, sample_dist):
    B, n_old = z_old.shape
    n_new = 0 if z_new is None else z_new.shape[1]
    z = torch.empty(B, n_old + n_new, dtype=torch.float32, device=z_old.device)
    mid = torch.empty_like(z)
    L.check(L.load().rnb_final_merge(L.ptr(z_old), n_old, L.ptr(z_new), n_new, B, float(sample_dist), L.ptr(z), L.ptr(mid),
                                     L.stream_ptr()), "final_merge")
    return z, mid


def composite_params(rays_o, rays_d, z, sdf, grad, albedo, lights, variance, cos_anneal_ratio, mode, sample_dist):
    """mode: 0 = render_rnb (no relu), 1 = render_rnb_warmup (relu), 2 = plain colour (shade = 1)"""
    B = z.shape[0]
    assert z.shape[1] == 128, "the compositing kernel is specialised for 128 samples per ray"
    p = L.Composite()
    p.n_rays = B
    p.rays_o, p.rays_d, p.z = L.ptr(rays_o), L.ptr(rays_d), L.ptr(z)
    p.sdf, p.grad, p.albedo = L.ptr(sdf), L.ptr(grad), L.ptr(albedo)
    nl = lights.shape[0]
    lights = _f32(lights).reshape(nl, -1, 3)
    p.lights, p.n_lights = L.ptr(lights), nl
    p.light_stride_l = lights.shape[1] * 3
    p.light_stride_ray = 3 if lights.shape[1] == B and B > 1 else (3 if lights.shape[1] == B else 0)
    if lights.shape[1] not in (1, B):
        raise RuntimeError(f"lights_dir must broadcast against [L,{B},1,3], got {tuple(lights.shape)}")
    if lights.shape[1] == 1:
        p.light_stride_ray = 0
    p.variance = L.ptr(variance)
    p.cos_anneal_ratio, p.warmup, p.sample_dist = float(cos_anneal_ratio), int(mode), float(sample_dist)
    p._keep = (lights, rays_o, rays_d, z, sdf, grad, albedo, variance)   # the struct only holds raw pointers
    p._device = z.device
    return p


def composite_fwd(p):
    B, nl, dev = p.n_rays, p.n_lights, p._device          # outputs live where the inputs live, not on the current device
    f32 = dict(dtype=torch.float32, device=dev)
    out = dict(color=torch.empty(nl, B, 3, **f32), weights=torch.empty(B, 128, **f32), cdf=torch.empty(B, 128, **f32),
               inside=torch.empty(B, 128, **f32), weight_sum=torch.empty(B, 1, **f32),
               weight_max=torch.empty(B, 1, **f32), eik_part=torch.empty(B, 2, **f32))
    p.color, p.weights, p.cdf, p.inside = (L.ptr(out[k]) for k in ("color", "weights", "cdf", "inside"))
    p.weight_sum, p.weight_max, p.eik_part = (L.ptr(out[k]) for k in ("weight_sum", "weight_max", "eik_part"))
    L.check(L.load().rnb_composite_fwd(C.byref(p), L.stream_ptr()), "composite_fwd")
    return out


def composite_bwd(p, d_color, d_weight_sum, d_eik, eik_den, want_albedo):
    B, dev = p.n_rays, d_color.device
    f32 = dict(dtype=torch.float32, device=dev)
    d_color = _f32(d_color)
    d_ws = _f32(d_weight_sum).view(-1) if d_weight_sum is not None else None
    out = dict(d_sdf=torch.empty(B * 128, **f32), d_grad=torch.empty(B * 128, 3, **f32),
               d_albedo=torch.empty(B * 128, 3, **f32) if want_albedo else None, d_var_part=torch.empty(B, **f32))
    p.d_color, p.d_weight_sum, p.d_eik, p.eik_den = L.ptr(d_color), L.ptr(d_ws), L.ptr(d_eik), L.ptr(eik_den)
    p.d_sdf, p.d_grad, p.d_albedo, p.d_var_part = (L.ptr(out[k]) for k in ("d_sdf", "d_grad", "d_albedo", "d_var_part"))
    L.check(L.load().rnb_composite_bwd(C.byref(p), L.stream_ptr()), "composite_bwd")
    return out


# ----------------------------------------------------------------------------- NeRF++ background
NERF_SHAPES = {"pts": [(256, 84)] + [(256, 256)] * 4 + [(256, 340)] + [(256, 256)] * 2,
               "feature_linear": (256, 256), "alpha_linear": (1, 256), "views_linears.0": (128, 283),
               "rgb_linear": (3, 128)}


class NerfPacked:
    """Packed fp16 operand images + fp32 side table of one NeRF state (D=8, W=256, skips=[4], multires 10/4)."""

    def __init__(self, nerf_module):
        lib = L.load()
        sd = {k: v.detach().float().contiguous() for k, v in nerf_module.state_dict().items()}
        Ws = [sd[f"pts_linears.{i}.weight"] for i in range(8)] if "pts_linears.7.weight" in sd else []
        ok = ([tuple(w.shape) for w in Ws] == NERF_SHAPES["pts"]
              and all(tuple(sd.get(k + ".weight", torch.empty(0)).shape) == NERF_SHAPES[k]
                      for k in ("feature_linear", "alpha_linear", "views_linears.0", "rgb_linear")))
        if not ok:
            raise RuntimeError("rnb_b200: the NeRF++ kernel is specialised for the shipped background field "
                               "(D=8, W=256, d_in=4, multires=10, multires_view=4, skips=[4], use_viewdirs=True)")
        dev = Ws[0].device
        L.require_cuda(Ws[0], "NeRF")
        bs = [sd[f"pts_linears.{i}.bias"] for i in range(8)]
        self.wblob = torch.empty(lib.rnb_nerf_wblob_bytes(), dtype=torch.uint8, device=dev)
        self.aux = torch.empty(lib.rnb_nerf_aux_floats(), dtype=torch.float32, device=dev)
        wp = (C.c_void_p * 8)(*[L.ptr(t) for t in Ws])
        bp = (C.c_void_p * 8)(*[L.ptr(t) for t in bs])
        rest = [sd[k] for k in ("feature_linear.weight", "feature_linear.bias", "alpha_linear.weight", "alpha_linear.bias",
                                "views_linears.0.weight", "views_linears.0.bias", "rgb_linear.weight", "rgb_linear.bias")]
        L.check(lib.rnb_nerf_pack(wp, bp, *[L.ptr(t) for t in rest], L.ptr(self.wblob), L.ptr(self.aux), L.stream_ptr()),
                "nerf_pack")
        self._keep = (Ws, bs, rest)


def nerf_fwd(packed: NerfPacked, pts=None, pts4=None, dirs=None):
    """-> (density [n], rgb [n,3]) raw head outputs; either ray samples `pts` (points_rays) or explicit pts4/dirs"""
    dev = packed.wblob.device
    if pts is None:
        pts4, dirs = _f32(pts4).view(-1, 4), _f32(dirs).view(-1, 3)
        L.require_cuda(pts4, "NeRF.forward")
        pts = _points(pts4.shape[0])
    n = pts.n_pts
    density = torch.empty(n, dtype=torch.float32, device=dev)
    rgb = torch.empty(n, 3, dtype=torch.float32, device=dev)
    L.check(L.load().rnb_nerf_fwd(C.byref(pts), L.ptr(pts4), L.ptr(dirs), L.ptr(packed.wblob), L.ptr(packed.aux),
                                  L.ptr(density), L.ptr(rgb), L.stream_ptr()), "nerf_fwd")
    return density, rgb


def composite_bg_fwd(rays_o, rays_d, z, sdf, grad, color_in, variance, cos_anneal_ratio, sample_dist, z_feed, bg_density,
                     bg_rgb):
    """render_core with background_alpha / background_sampled_color (reference models/renderer.py:194-285), forward"""
    B, n_tot = z_feed.shape
    assert z.shape[1] == 128, "the compositing kernels are specialised for 128 SDF samples per ray"
    dev = z.device
    f32 = dict(dtype=torch.float32, device=dev)
    out = dict(color=torch.empty(B, 3, **f32), weights=torch.empty(B, n_tot, **f32), cdf=torch.empty(B, 128, **f32),
               inside=torch.empty(B, 128, **f32), weight_sum=torch.empty(B, 1, **f32),
               weight_max=torch.empty(B, 1, **f32), eik_part=torch.empty(B, 2, **f32))
    p = L.CompositeBg()
    p.n_rays = B
    keep = [_f32(t) for t in (rays_o, rays_d, z, sdf, grad, color_in, variance, z_feed, bg_density, bg_rgb)]
    (p.rays_o, p.rays_d, p.z, p.sdf, p.grad, p.color_in, p.variance, p.z_feed, p.bg_density,
     p.bg_rgb) = (L.ptr(t) for t in keep)
    p.cos_anneal_ratio, p.sample_dist, p.n_outside = float(cos_anneal_ratio), float(sample_dist), n_tot - 128
    p.color, p.weights, p.cdf, p.inside = (L.ptr(out[k]) for k in ("color", "weights", "cdf", "inside"))
    p.weight_sum, p.weight_max, p.eik_part = (L.ptr(out[k]) for k in ("weight_sum", "weight_max", "eik_part"))
    L.check(L.load().rnb_composite_bg_fwd(C.byref(p), L.stream_ptr()), "composite_bg_fwd")
    return out


def stream_from_rowmajor(x, cols):
    """fp32 [n, cols] row-major -> fp16 stream image [Npad/64][cols/8][64][8] (the layout the chain kernels read)"""
    x = _f32(x).view(-1, cols)
    L.require_cuda(x, "stream_from_rowmajor")
    lib = L.load()
    n = x.shape[0]
    buf = torch.empty(lib.rnb_stream_bytes(n, cols), dtype=torch.uint8, device=x.device)
    L.check(lib.rnb_stream_from_rowmajor(L.ptr(x), n, cols, L.ptr(buf), L.stream_ptr()), "stream_from_rowmajor")
    return buf


def stream_to_rowmajor(buf, n_pts, cols, dtype=torch.float16):
    """Decode a stream image [Npad/64][cols/8][64][8] into [n_pts, cols] (tests / debugging)."""
    t = buf.view(dtype).view(-1, cols // 8, 64, 8).permute(0, 2, 1, 3).reshape(-1, cols)
    return t[:n_pts]
