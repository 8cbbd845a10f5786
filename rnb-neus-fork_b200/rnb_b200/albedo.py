"""Albedo network (RenderingNetwork, mode 'no_view_dir') forward / backward on the kernels of csrc/albedo.cu."""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as L

ALBEDO_SHAPES = [(256, 310), (256,), (256, 256), (256,), (3, 256), (3,)]


class AlbedoCtx:
    pass


def pack(flat, device):
    lib = L.load()
    if [tuple(t.shape) for t in flat] != ALBEDO_SHAPES:
        raise RuntimeError("rnb_b200: the albedo kernels are specialised for the shipped rendering_network "
                           f"(d_feature=256, d_hidden=256, n_layers=2, multires_view=4, mode=no_view_dir); got "
                           f"{[tuple(t.shape) for t in flat]}")
    blob = torch.empty(lib.rnb_albedo_wblob_bytes(), dtype=torch.uint8, device=device)
    aux = torch.empty(lib.rnb_albedo_aux_floats(), dtype=torch.float32, device=device)
    keep = [t.detach().float().contiguous() for t in flat]
    L.check(lib.rnb_albedo_pack(*[L.ptr(t) for t in keep], L.ptr(blob), L.ptr(aux), L.stream_ptr()), "albedo_pack")
    return blob, aux, keep


def forward(flat, pts, normals, sdf_streams, for_backward=True):
    """-> AlbedoCtx with .albedo [n,3] and (for_backward) the streams kept for the backward."""
    lib = L.load()
    dev = normals.device
    n = pts.n_pts
    ctx = AlbedoCtx()
    ctx.blob, ctx.aux, ctx.keep = pack(flat, dev)
    ctx.pts, ctx.normals, ctx.feat = pts, normals, sdf_streams.feat
    u8 = dict(dtype=torch.uint8, device=dev)
    ctx.st_pe = torch.empty(lib.rnb_stream_bytes(n, 64), **u8) if for_backward else None
    ctx.st_h0 = torch.empty(lib.rnb_stream_bytes(n, 256), **u8) if for_backward else None
    ctx.st_h1 = torch.empty(lib.rnb_stream_bytes(n, 256), **u8) if for_backward else None
    ctx.albedo = torch.empty(n, 3, dtype=torch.float32, device=dev)
    L.check(lib.rnb_albedo_fwd(C.byref(pts), L.ptr(normals), L.ptr(ctx.feat), L.ptr(ctx.blob), L.ptr(ctx.aux),
                               L.ptr(ctx.albedo), L.ptr(ctx.st_pe), L.ptr(ctx.st_h0), L.ptr(ctx.st_h1), L.stream_ptr()),
            "albedo_fwd")
    return ctx


def backward(ctx, d_albedo, want_fp32=False):
    """-> (d_normal [n,3], d_feat, [dW0, db0, dW1, db1, dW2, db2]); d_feat is (fp16 stream, meta float[2]) for
    kernels.sdf_bwd(d_feat16=...), or the fp32 [n,256] tensor when want_fp32"""
    lib = L.load()
    dev = ctx.albedo.device
    n = ctx.pts.n_pts
    d_albedo = d_albedo.detach().float().contiguous().view(-1, 3)
    scratch = torch.empty(lib.rnb_albedo_bwd_scratch_bytes(n), dtype=torch.uint8, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    d_normal = torch.empty(n, 3, **f32)
    d_feat = torch.empty(n, 256, **f32) if want_fp32 else None
    d_feat16 = None if want_fp32 else torch.empty(lib.rnb_stream_bytes(n, 256), dtype=torch.uint8, device=dev)
    meta = None if want_fp32 else torch.empty(2, **f32)
    grads = [torch.empty(s, **f32) for s in ALBEDO_SHAPES]
    L.check(lib.rnb_albedo_bwd(C.byref(ctx.pts), L.ptr(ctx.normals), L.ptr(ctx.albedo), L.ptr(d_albedo), L.ptr(ctx.feat),
                               L.ptr(ctx.st_pe), L.ptr(ctx.st_h0), L.ptr(ctx.st_h1), L.ptr(ctx.blob), L.ptr(ctx.aux),
                               L.ptr(scratch), L.ptr(d_normal), L.ptr(d_feat), L.ptr(d_feat16), L.ptr(meta),
                               *[L.ptr(g) for g in grads], L.stream_ptr()),
            "albedo_bwd")
    return d_normal, (d_feat if want_fp32 else (d_feat16, meta)), grads
