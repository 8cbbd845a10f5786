"""Autograd glue between the reference-shaped modules (models/fields.py, models/renderer.py) and the kernels.

Host logic only: which kernel runs when, which buffers are kept for the backward, and how the cotangents flow
(composite adjoint -> albedo net backward -> SDF double-backward).  Weight-norm stays in torch on parameter-sized
tensors (W = g v/|v| via torch._weight_norm), so autograd turns the kernels' dW into (dg, dv) for free.
"""
from __future__ import annotations

import torch

from . import kernels as K
from . import lib as L

FINE_SAMPLES = 128


def _params_of(module):
    """list(module.parameters()), cached on the module (walking the module tree every step showed up in the host profile of
    the 512-ray iteration); parameters are never re-created on the hot path, only updated in place."""
    ps = getattr(module, "_rnb_params", None)
    if ps is None:
        ps = list(module.parameters())
        module._rnb_params = ps
    return ps


def _sdf_wb(sdf_module):
    eff = sdf_module.effective_weights()
    if len(eff) != 9:
        raise RuntimeError("rnb_b200: the sm_100a SDF kernels are specialised for n_layers=8 (9 linears)")
    if getattr(sdf_module, "multires", 6) != 6 or tuple(sdf_module.skip_in) != (4,) or float(sdf_module.scale) != 1.0:
        raise RuntimeError("rnb_b200: SDF kernels support multires=6, skip_in=[4], scale=1.0 (the shipped confs)")
    flat = []
    for W, b in eff:
        flat += [W, b]
    return flat


def _pack_sdf(flat, device):
    pk = K.SdfPacked(device)
    pk.pack(flat[0::2], flat[1::2])
    return pk


def packed_sdf_nograd(sdf_module):
    """Packed operands for no_grad evaluation, cached until a parameter changes (version counters)."""
    params = _params_of(sdf_module)
    key = tuple((p.data_ptr(), p._version) for p in params)
    cache = getattr(sdf_module, "_rnb_packed", None)
    if torch.cuda.is_current_stream_capturing():
        # CUDA-graph capture (graph_step.py): the pack kernel must be PART of the graph, or replays after an
        # optimiser step would sample with the weights of capture time
        with torch.no_grad():
            flat = _sdf_wb(sdf_module)
            return _pack_sdf(flat, flat[0].device)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            flat = _sdf_wb(sdf_module)
            L.require_cuda(flat[0], "SDFNetwork")
            cache = (key, _pack_sdf(flat, flat[0].device))
        sdf_module._rnb_packed = cache
    return cache[1]


def fold_and_pack(sdf_module):
    """One weight-norm fold + one pack for a whole training step: -> (flat, pk).  `flat` are the effective weights WITH
    their autograd graph (the fine pass differentiates through them), `pk` the packed operands built from the same
    tensors; the no_grad sampling pass and the fine pass both use `pk`, instead of one fold + pack each
    (the reference re-folds inside every SDFNetwork.forward call, models/fields.py:72-74)."""
    flat = _sdf_wb(sdf_module)
    L.require_cuda(flat[0], "SDFNetwork")
    with torch.no_grad():
        pk = _pack_sdf(flat, flat[0].device)
    if not torch.cuda.is_current_stream_capturing():
        params = _params_of(sdf_module)
        sdf_module._rnb_packed = (tuple((p.data_ptr(), p._version) for p in params), pk)
    return flat, pk


# ------------------------------------------------------------------------------------------ module-level API
def sdf_only(sdf_module, x):
    """SDFNetwork.sdf under no_grad -> [N,1]"""
    L.require_cuda(x, "SDFNetwork.sdf")
    pk = packed_sdf_nograd(sdf_module)
    return K.sdf_fwd(pk, K.points_explicit(x.reshape(-1, 3))).view(-1, 1)


class _SdfEval(torch.autograd.Function):
    """(out [N,257], grad [N,3]) of explicit points; differentiable w.r.t. the effective weights."""

    @staticmethod
    def forward(ctx, x, *flat):
        pk = _pack_sdf(flat, x.device)
        pts = K.points_explicit(x)
        sdf, grad, full, streams = K.sdf_fwd_grad(pk, pts, want_full=True)
        ctx.pk, ctx.pts, ctx.streams = pk, pts, streams
        return full, grad

    @staticmethod
    def backward(ctx, d_full, d_grad):
        n = ctx.pts.n_pts
        dev = ctx.pk.wblob.device
        if d_full is None:
            d_full = torch.zeros(n, 257, device=dev)
        if d_grad is None:
            d_grad = torch.zeros(n, 3, device=dev)
        d_sdf = d_full[:, 0].contiguous()
        d_feat = d_full[:, 1:].contiguous()
        dWs, dbs, _ = K.sdf_bwd(ctx.pk, ctx.pts, ctx.streams, d_sdf, d_grad, d_feat)
        out = [None]
        for W, b in zip(dWs, dbs):
            out += [W, b]
        return tuple(out)


def sdf_forward(sdf_module, x, want_grad):
    """-> (out [N,257], grad [N,3] or None)"""
    L.require_cuda(x, "SDFNetwork.forward")
    x2 = x.detach().reshape(-1, 3).float().contiguous()
    flat = _sdf_wb(sdf_module)
    needs = torch.is_grad_enabled() and any(t.requires_grad for t in flat)
    if needs:
        full, grad = _SdfEval.apply(x2, *flat)
    else:
        with torch.no_grad():
            pk = packed_sdf_nograd(sdf_module)
            _, grad, full, _ = K.sdf_fwd_grad(pk, K.points_explicit(x2), want_full=True)
    return full, (grad if want_grad else None)


def albedo_forward(color_module, points, normals, view_dirs, feature_vectors):
    """Stand-alone RenderingNetwork.forward, mode 'no_view_dir' (reference models/fields.py:177-215; called directly by
    validate_mesh_texture, exp_runner.py:613-615).  Forward only: its reference caller detaches the result."""
    L.require_cuda(points, "RenderingNetwork.forward")
    if getattr(color_module, "mode", "no_view_dir") != "no_view_dir":
        raise RuntimeError("rnb_b200: the albedo kernels implement mode='no_view_dir' (the shipped confs)")
    from . import albedo as A
    with torch.no_grad():
        flat = []
        for W, b in color_module.effective_weights():
            flat += [W, b]
        pts = K.points_explicit(points.reshape(-1, 3))

        class _S:
            pass
        st = _S()
        st.feat = K.stream_from_rowmajor(feature_vectors.reshape(-1, 256), 256)
        nrm = normals.detach().reshape(-1, 3).float().contiguous()
        out = A.forward(flat, pts, nrm, st).albedo
    return out


def packed_nerf(nerf_module):
    """Packed NeRF operands, cached until a parameter changes (version counters)."""
    params = _params_of(nerf_module)
    key = tuple((p.data_ptr(), p._version) for p in params)
    cache = getattr(nerf_module, "_rnb_packed", None)
    if torch.cuda.is_current_stream_capturing():
        with torch.no_grad():
            return K.NerfPacked(nerf_module)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            cache = (key, K.NerfPacked(nerf_module))
        nerf_module._rnb_packed = cache
    return cache[1]


def nerf_forward(nerf_module, input_pts, input_views):
    """NeRF.forward(input_pts [n,4], input_views [n,3]) -> (alpha [n,1], rgb [n,3]) (reference models/fields.py:281-314).
    Forward only (the background model is reached only through render() -> render_novel_image upstream)."""
    L.require_cuda(input_pts, "NeRF.forward")
    with torch.no_grad():
        dens, rgb = K.nerf_fwd(packed_nerf(nerf_module), pts4=input_pts.reshape(-1, 4), dirs=input_views.reshape(-1, 3))
    return dens.view(-1, 1), rgb


@torch.no_grad()
def render_with_background(renderer, rays_o, rays_d, z_vals, mid_z, z_outside, cos_anneal_ratio, sample_dist):
    """render() with n_outside > 0 (reference models/renderer.py:609-631 + render_core_outside :93-130 + the blend in
    render_core :255-260), forward only.  -> dict of the compositing kernel's outputs + gradients."""
    o = rays_o.detach().float().contiguous()
    d = rays_d.detach().float().contiguous()
    B = o.shape[0]
    # SDF pass on the 128 hierarchical samples
    pk = packed_sdf_nograd(renderer.sdf_network)
    pts = K.points_rays(o, d, mid_z)
    sdf, grad, _, streams = K.sdf_fwd_grad(pk, pts)
    from . import albedo as A
    col = []
    for W, b in renderer.color_network.effective_weights():
        col += [W, b]
    color_in = A.forward(col, pts, grad, streams).albedo
    # background pass on sort(cat[z_vals, z_vals_outside]) (renderer.py:610-612); both lists are sorted -> rank merge
    z_feed, mid_feed = K.final_merge(z_vals, z_outside.float().contiguous(), sample_dist)
    dens, rgb = K.nerf_fwd(packed_nerf(renderer.nerf), pts=K.points_rays(o, d, mid_feed))
    var = renderer.deviation_network.variance.detach().float().reshape(1).contiguous()
    out = K.composite_bg_fwd(o, d, z_vals, sdf, grad, color_in, var, cos_anneal_ratio, sample_dist, z_feed, dens, rgb)
    out["gradients"] = grad.view(B, FINE_SAMPLES, 3)
    return out


# ------------------------------------------------------------------------------------------ sampling
@torch.no_grad()
def hierarchical_sample(sdf_module, rays_o, rays_d, near, far, t_rand, n_samples, n_importance, up_sample_steps, pk=None):
    """The no_grad block of render*/render_rnb* (reference models/renderer.py:829-880): -> (z_vals, mid_z_vals).
    pk: packed operands of this step (fold_and_pack), else the cached no_grad pack."""
    if pk is None:
        pk = packed_sdf_nograd(sdf_module)
    o = rays_o.detach().float().contiguous()
    d = rays_d.detach().float().contiguous()
    B = o.shape[0]
    z = K.coarse_z(near, far, t_rand, n_samples)
    last = None
    if n_importance > 0:
        sdf = K.sdf_fwd(pk, K.points_rays(o, d, z)).view(B, n_samples)
        n_new = n_importance // up_sample_steps
        pend_z = pend_s = None
        for i in range(up_sample_steps):
            z_new, z, sdf, _, _ = K.upsample_step(o, d, z, sdf, 64 * 2 ** i, n_new, pend_z, pend_s)
            if i + 1 < up_sample_steps:
                pend_z = z_new
                pend_s = K.sdf_fwd(pk, K.points_rays(o, d, z_new)).view(B, n_new)
            else:
                last = z_new
    return K.final_merge(z, last, 2.0 / n_samples)


# ------------------------------------------------------------------------------------------ fine pass
class _RnbFine(torch.autograd.Function):
    """render_core_mvps + RNb shading (reference models/renderer.py:466-554, 904-918, 1008-1017) as one node.

    inputs : meta dict, rays_o, rays_d, z_vals, mid_z, lights, variance, 18 SDF tensors (W_l, b_l), [6 colour tensors]
    outputs: color [L,B,3], weight_sum [B,1], gradient_error [] (differentiable);
             weights, cdf, inside_sphere, weight_max, gradients [B,128,3], sdf [B*128,1] (not differentiable)
    """

    @staticmethod
    def forward(ctx, meta, rays_o, rays_d, z_vals, mid_z, lights, variance, *wb):
        dev = z_vals.device
        B = z_vals.shape[0]
        sdf_flat = wb[:18]
        col_flat = wb[18:]
        use_albedo = meta["use_albedo"]
        pk = meta.get("pk") or _pack_sdf(sdf_flat, dev)
        pts = K.points_rays(rays_o, rays_d, mid_z)
        sdf, grad, _, streams = K.sdf_fwd_grad(pk, pts)
        albedo = None
        if use_albedo:
            from . import albedo as A
            actx = A.forward(col_flat, pts, grad, streams)
            albedo = actx.albedo
            ctx.actx = actx
        var = variance.detach().float().reshape(1).contiguous()
        cp = K.composite_params(rays_o, rays_d, z_vals, sdf, grad, albedo, lights, var, meta["cos_anneal_ratio"],
                                meta["mode"], meta["sample_dist"])
        out = K.composite_fwd(cp)
        eik = out["eik_part"].sum(0)
        if meta.get("dp_group") is not None:
            # exact global-batch normaliser (parallel.ExactBatch): the eikonal mean runs over the points of ALL ranks --
            # numerator and count (reference models/renderer.py:540) are summed over the group, two floats in one all-reduce.
            # The backward differentiates the global ratio through this rank's points only: exactly its share.
            import torch.distributed as dist
            dist.all_reduce(eik, op=dist.ReduceOp.SUM, group=meta["dp_group"] if meta["dp_group"] is not True else None)
        grad_err = eik[0] / (eik[1] + 1e-5)
        ctx.pk, ctx.pts, ctx.streams, ctx.cp = pk, pts, streams, cp
        ctx.eik_den = eik[1:2].contiguous()
        ctx.use_albedo = use_albedo
        ctx.n_col = len(col_flat)
        gradients = grad.view(B, FINE_SAMPLES, 3)
        sampled_albedo = albedo.view(B, FINE_SAMPLES, 3) if albedo is not None else None
        nd = [out["weights"], out["cdf"], out["inside"], out["weight_max"], gradients, sdf]
        if sampled_albedo is not None:
            nd.append(sampled_albedo)
        ctx.mark_non_differentiable(*nd)
        return (out["color"], out["weight_sum"], grad_err, out["weights"], out["cdf"], out["inside"], out["weight_max"],
                gradients, sdf, sampled_albedo)

    @staticmethod
    def backward(ctx, d_color, d_wsum, d_eik, *unused):
        cp = ctx.cp
        dev = ctx.pk.wblob.device
        B, nl = cp.n_rays, cp.n_lights
        if d_color is None:
            d_color = torch.zeros(nl, B, 3, device=dev)
        if d_eik is None:
            d_eik = torch.zeros((), device=dev)
        d_eik = d_eik.detach().float().reshape(1).contiguous()
        bw = K.composite_bwd(cp, d_color, d_wsum, d_eik, ctx.eik_den, ctx.use_albedo)
        d_grad = bw["d_grad"]
        d_feat16 = None
        col_grads = [None] * ctx.n_col
        if ctx.use_albedo:
            from . import albedo as A
            d_normal, d_feat16, col_grads = A.backward(ctx.actx, bw["d_albedo"])
            d_grad = d_grad + d_normal
        dWs, dbs, _ = K.sdf_bwd(ctx.pk, ctx.pts, ctx.streams, bw["d_sdf"], d_grad, None, d_feat16=d_feat16)
        d_var = bw["d_var_part"].sum().reshape(())
        grads = [None, None, None, None, None, None, d_var]
        for W, b in zip(dWs, dbs):
            grads += [W, b]
        grads += list(col_grads)
        return tuple(grads)


@torch.no_grad()
def _rnb_fine_inference(sdf_module, color_module, variance, o, d, z_vals, mid_z, lights, meta):
    """Forward-only fine pass (validate_image / render_novel_image under no_grad, exp_runner.py:389-470, 519-558):
    nothing is kept for a backward, so the kernels skip the streams only the backward reads (K2 writes 4.7 instead of
    8.8 KB/point, the albedo net 0 instead of 1.1) and the cached packed weights are reused."""
    B = z_vals.shape[0]
    pk = packed_sdf_nograd(sdf_module)
    pts = K.points_rays(o, d, mid_z)
    sdf, grad, _, streams = K.sdf_fwd_grad(pk, pts, for_backward=False)
    albedo = None
    if meta["use_albedo"]:
        from . import albedo as A
        col = []
        for W, b in color_module.effective_weights():
            col += [W, b]
        albedo = A.forward(col, pts, grad, streams, for_backward=False).albedo
    var = variance.detach().float().reshape(1).contiguous()
    cp = K.composite_params(o, d, z_vals, sdf, grad, albedo, lights, var, meta["cos_anneal_ratio"], meta["mode"],
                            meta["sample_dist"])
    out = K.composite_fwd(cp)
    eik = out["eik_part"].sum(0)
    return (out["color"], out["weight_sum"], eik[0] / (eik[1] + 1e-5), out["weights"], out["cdf"], out["inside"],
            out["weight_max"], grad.view(B, FINE_SAMPLES, 3), sdf,
            albedo.view(B, FINE_SAMPLES, 3) if albedo is not None else None)


def rnb_fine(sdf_module, color_module, variance, rays_o, rays_d, z_vals, mid_z, lights, cos_anneal_ratio, mode,
             use_albedo, sample_dist, folded=None, dp_group=None):
    """mode 0: render_rnb, 1: render_rnb_warmup, 2: plain colour (render).  folded: (flat, pk) of fold_and_pack."""
    if z_vals.shape[1] != FINE_SAMPLES:
        raise RuntimeError(f"rnb_b200: the fine pass is specialised for n_samples + n_importance = {FINE_SAMPLES} "
                           f"(got {z_vals.shape[1]})")
    flat, pk = folded if folded is not None else (_sdf_wb(sdf_module), None)
    col = []
    if use_albedo:
        for W, b in color_module.effective_weights():
            col += [W, b]
    meta = dict(use_albedo=use_albedo, cos_anneal_ratio=float(cos_anneal_ratio), mode=int(mode),
                sample_dist=float(sample_dist), pk=pk, dp_group=dp_group)
    o = rays_o.detach().float().contiguous()
    d = rays_d.detach().float().contiguous()
    needs_grad = torch.is_grad_enabled() and (variance.requires_grad or any(t.requires_grad for t in flat + col))
    if not needs_grad:
        return _rnb_fine_inference(sdf_module, color_module, variance, o, d, z_vals, mid_z, lights.detach(), meta)
    return _RnbFine.apply(meta, o, d, z_vals, mid_z, lights.detach(), variance, *flat, *col)
