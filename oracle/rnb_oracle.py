"""CPU oracle for the RNb-NeuS hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy (float64) restatement of the reference's algorithm for the
`train_rnb` ray-batch step and the SDF grid query.  It is the checker the CUDA
path is compared against.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it; the product
(`rnb-neus-fork_b200/`) never does and fails loudly without its CUDA library.

Pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against the reference's own PyTorch code executed in the build
container: `oracle/gen_golden.py` imports `/root/reference/models/*` and writes
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every oracle function
against those fixtures (and, when /root/reference is present, against the live
reference).  Everything the reference gets from autograd (d sdf/dx, the
double-backward, the compositing adjoint) is restated analytically here.

All citations are file:line relative to /root/reference.
"""
from __future__ import annotations

import numpy as np

F64 = np.float64
SQRT2 = np.sqrt(2.0)

# --------------------------------------------------------------------------
# A1  positional encoding                      models/embedder.py:12-55, 58-74
# --------------------------------------------------------------------------


def embed(x, multires):
    """[N,d] -> [N, d*(1+2*multires)]; column order x | sin f0 | cos f0 | sin f1 ...
    with freq = 2**linspace(0, L-1, L) (models/embedder.py:34, 39-45)."""
    x = np.asarray(x, F64)
    out = [x]
    for k in range(multires):
        f = 2.0 ** k
        out.append(np.sin(x * f))
        out.append(np.cos(x * f))
    return np.concatenate(out, -1)


def embed_vjp(x, de, multires):
    """J_e^T de : [N, d*(1+2L)] -> [N,d]   (what autograd does to models/embedder.py:53-55)."""
    x = np.asarray(x, F64)
    d = x.shape[-1]
    g = de[:, :d].copy()
    for k in range(multires):
        f = 2.0 ** k
        s = de[:, d + 2 * d * k: d + 2 * d * k + d]
        c = de[:, d + 2 * d * k + d: d + 2 * d * k + 2 * d]
        g += f * (np.cos(f * x) * s - np.sin(f * x) * c)
    return g


def embed_jvp(x, gbar, multires):
    """J_e gbar : [N,d] -> [N, d*(1+2L)]."""
    x = np.asarray(x, F64)
    out = [gbar]
    for k in range(multires):
        f = 2.0 ** k
        out.append(f * np.cos(f * x) * gbar)
        out.append(-f * np.sin(f * x) * gbar)
    return np.concatenate(out, -1)


# --------------------------------------------------------------------------
# softplus(beta=100) with the ATen threshold   models/fields.py:80 (nn.Softplus)
# --------------------------------------------------------------------------
BETA = 100.0
THRESH = 20.0


def softplus(z):
    bz = BETA * z
    safe = np.minimum(bz, THRESH)
    return np.where(bz > THRESH, z, np.log1p(np.exp(safe)) / BETA)


def softplus_d1(z):
    bz = BETA * z
    safe = np.minimum(bz, THRESH)
    return np.where(bz > THRESH, 1.0, 1.0 / (1.0 + np.exp(-safe)))


def softplus_d2(z):
    s = softplus_d1(z)
    return np.where(BETA * z > THRESH, 0.0, BETA * s * (1.0 - s))


def sigmoid(x):
    x = np.asarray(x, F64)
    return np.where(x >= 0, 1.0 / (1.0 + np.exp(-np.abs(x))), np.exp(-np.abs(x)) / (1.0 + np.exp(-np.abs(x))))


# --------------------------------------------------------------------------
# weight norm                                   models/fields.py:72-74
# --------------------------------------------------------------------------


def weight_norm_fold(g, v):
    """W = g * v / ||v||_row   (legacy nn.utils.weight_norm, dim=0)."""
    g = np.asarray(g, F64)
    v = np.asarray(v, F64)
    n = np.sqrt((v * v).sum(1, keepdims=True))
    return g * v / n


def weight_norm_vjp(g, v, dW):
    """(dg, dv) from dW  (SURVEY 8a' K3 step 4)."""
    g = np.asarray(g, F64)
    v = np.asarray(v, F64)
    n = np.sqrt((v * v).sum(1, keepdims=True))
    dot = (dW * v).sum(1, keepdims=True)
    dg = dot / n
    dv = g / n * (dW - v * dot / (n * n))
    return dg, dv


# --------------------------------------------------------------------------
# parameter containers: plain dicts of numpy arrays keyed like the state_dict
#   sdf:   lin{l}.weight_g [out,1], lin{l}.weight_v [out,in], lin{l}.bias [out]
#   color: same keys, lin0..lin2
# --------------------------------------------------------------------------


def sdf_effective(sd, n_lin=9):
    Ws = [weight_norm_fold(sd[f"lin{l}.weight_g"], sd[f"lin{l}.weight_v"]) for l in range(n_lin)]
    bs = [np.asarray(sd[f"lin{l}.bias"], F64) for l in range(n_lin)]
    return Ws, bs


# --------------------------------------------------------------------------
# A2/A3  SDFNetwork.forward / .sdf              models/fields.py:82-108
# --------------------------------------------------------------------------


def sdf_forward(Ws, bs, x, multires=6, skip_in=(4,), scale=1.0, keep=False):
    """Returns out [N, d_out] (col 0 = sdf / scale, cols 1: = features).
    With keep=True also returns the per-layer (in_l, z_l) lists for the adjoints."""
    x = np.asarray(x, F64)
    inputs = x * scale
    e = embed(inputs, multires) if multires > 0 else inputs
    a = e
    n_lin = len(Ws)
    ins, zs = [], []
    for l in range(n_lin):
        if l in skip_in:
            a = np.concatenate([a, e], 1) / SQRT2          # models/fields.py:94-96
        z = a @ Ws[l].T + bs[l]
        ins.append(a)
        zs.append(z)
        a = softplus(z) if l < n_lin - 1 else z             # models/fields.py:100-102
    out = np.concatenate([a[:, :1] / scale, a[:, 1:]], -1)  # models/fields.py:104
    if keep:
        return out, ins, zs, e
    return out


def sdf_only(Ws, bs, x, **kw):
    return sdf_forward(Ws, bs, x, **kw)[:, :1]


# --------------------------------------------------------------------------
# A4  SDFNetwork.gradient, analytic dx-chain    models/fields.py:114-127
#     (SURVEY 8a' K2; scale == 1 as in every shipped conf)
# --------------------------------------------------------------------------


def sdf_gradient(Ws, bs, x, multires=6, skip_in=(4,), keep=False):
    x = np.asarray(x, F64)
    out, ins, zs, e = sdf_forward(Ws, bs, x, multires, skip_in, keep=True)
    n_lin = len(Ws)
    N = x.shape[0]
    d_e = e.shape[1]
    ua = np.broadcast_to(Ws[n_lin - 1][0:1, :], (N, Ws[n_lin - 1].shape[1])).copy()  # seed: d z8[0] / d a7
    de_skip = np.zeros((N, d_e))
    uas = [None] * (n_lin - 1)
    ws = [None] * (n_lin - 1)
    for l in range(n_lin - 2, -1, -1):
        s = softplus_d1(zs[l])
        w = s * ua
        uas[l], ws[l] = ua, w
        uin = w @ Ws[l]
        if l in skip_in:
            uin = uin / SQRT2
            n_prev = uin.shape[1] - d_e
            de_skip = de_skip + uin[:, n_prev:]
            ua = uin[:, :n_prev]
        elif l > 0:
            ua = uin
    de = uin + de_skip
    g = embed_vjp(x, de, multires)
    if keep:
        return out, g, ins, zs, e, uas, ws
    return g


# --------------------------------------------------------------------------
# A5  double-backward of (sdf, feat, grad)      exp_runner.py:261 via autograd
#     (SURVEY 8a' K3)  -> grads w.r.t. the effective W_l, b_l
# --------------------------------------------------------------------------


def sdf_backward(Ws, bs, x, ybar, gbar, multires=6, skip_in=(4,)):
    """ybar [N, d_out] cotangent of forward output, gbar [N,3] cotangent of gradient().
    Returns (dWs, dbs) for the effective (weight-norm folded) weights."""
    x = np.asarray(x, F64)
    ybar = np.asarray(ybar, F64)
    gbar = np.asarray(gbar, F64)
    out, g, ins, zs, e, uas, ws = sdf_gradient(Ws, bs, x, multires, skip_in, keep=True)
    n_lin = len(Ws)
    dWs = [np.zeros_like(W) for W in Ws]
    dbs = [np.zeros_like(b) for b in bs]
    # step 1+2: adjoint of the dx-chain, runs forward through the layers
    ebar = embed_jvp(x, gbar, multires)
    uin_bar = ebar
    z2 = [None] * (n_lin - 1)
    for l in range(n_lin - 1):
        if l in skip_in:
            uin_bar = np.concatenate([uin_bar, ebar], 1) / SQRT2
        wbar = uin_bar @ Ws[l].T
        dWs[l] += ws[l].T @ uin_bar
        z2[l] = softplus_d2(zs[l]) * uas[l] * wbar
        ua_bar = softplus_d1(zs[l]) * wbar
        uin_bar = ua_bar
    dWs[n_lin - 1][0, :] += ua_bar.sum(0)
    # step 3: ordinary backward
    zbar = ybar
    for l in range(n_lin - 1, -1, -1):
        dWs[l] += zbar.T @ ins[l]
        dbs[l] += zbar.sum(0)
        if l == 0:
            break
        inbar = zbar @ Ws[l]
        if l in skip_in:
            n_prev = Ws[l - 1].shape[0]
            abar = inbar[:, :n_prev] / SQRT2
        else:
            abar = inbar
        zbar = softplus_d1(zs[l - 1]) * abar + z2[l - 1]
    return dWs, dbs


# --------------------------------------------------------------------------
# A6  RenderingNetwork.forward (mode no_view_dir) models/fields.py:177-215
# --------------------------------------------------------------------------


def color_forward(Ws, bs, points, normals, feat, multires_view=4, keep=False):
    pe = embed(points, multires_view)
    ne = embed(normals, multires_view)
    h = np.concatenate([pe, ne, np.asarray(feat, F64)], -1)     # models/fields.py:192
    ins, zs = [], []
    n_lin = len(Ws)
    for l in range(n_lin):
        z = h @ Ws[l].T + bs[l]
        ins.append(h)
        zs.append(z)
        h = np.maximum(z, 0.0) if l < n_lin - 1 else z            # models/fields.py:207-210
    out = sigmoid(h)                                               # models/fields.py:213
    if keep:
        return out, ins, zs
    return out


def color_backward(Ws, bs, points, normals, feat, d_out, multires_view=4, relu_masks=None):
    """Returns dWs, dbs, d_normals [N,3], d_feat [N,F].  (points carry no gradient:
    sample positions are detached, models/renderer.py:175, 862.)
    relu_masks (optional, list per hidden layer): use these ReLU activity masks instead of the oracle's own --
    a reduced-precision forward flips the sign of the few pre-activations that sit within its rounding error of
    zero, and a flipped unit changes that point's cotangent by 100 %; tests that check the backward ARITHMETIC
    feed the kernel's masks."""
    out, ins, zs = color_forward(Ws, bs, points, normals, feat, multires_view, keep=True)
    n_lin = len(Ws)
    dWs = [None] * n_lin
    dbs = [None] * n_lin
    zbar = np.asarray(d_out, F64) * out * (1.0 - out)
    for l in range(n_lin - 1, -1, -1):
        dWs[l] = zbar.T @ ins[l]
        dbs[l] = zbar.sum(0)
        hbar = zbar @ Ws[l]
        if l > 0:
            mask = (zs[l - 1] > 0) if relu_masks is None else relu_masks[l - 1]
            zbar = hbar * mask
    d_pe = 3 * (1 + 2 * multires_view)
    d_ne = hbar[:, d_pe:2 * d_pe]
    d_feat = hbar[:, 2 * d_pe:]
    d_normals = embed_vjp(normals, d_ne, multires_view)
    return dWs, dbs, d_normals, d_feat


# --------------------------------------------------------------------------
# A10  sample_pdf (det=True)                     models/renderer.py:39-69
# --------------------------------------------------------------------------


def searchsorted_right(cdf, u):
    """torch.searchsorted(cdf, u, right=True): first index i with cdf[i] > u. Row-wise."""
    B = cdf.shape[0]
    out = np.empty(u.shape, np.int64)
    for b in range(B):
        out[b] = np.searchsorted(cdf[b], u[b], side="right")
    return out


def sample_pdf_det(bins, weights, n_samples, dtype=np.float32, cdf_override=None):
    """Computed in `dtype` (float32 by default, like the reference) so that the
    searchsorted indices can be compared bit-exactly.  Returns (samples, inds, cdf)."""
    bins = np.asarray(bins, dtype)
    w = np.asarray(weights, dtype) + dtype(1e-5)
    pdf = w / w.sum(-1, keepdims=True, dtype=dtype)
    cdf = np.cumsum(pdf, -1, dtype=dtype)
    cdf = np.concatenate([np.zeros_like(cdf[:, :1]), cdf], -1)
    if cdf_override is not None:
        cdf = np.asarray(cdf_override, dtype)
    u = np.linspace(0.5 / n_samples, 1.0 - 0.5 / n_samples, n_samples).astype(dtype)
    u = np.broadcast_to(u, (cdf.shape[0], n_samples))
    inds = searchsorted_right(cdf, u)
    below = np.maximum(inds - 1, 0)
    above = np.minimum(inds, cdf.shape[-1] - 1)
    cdf_b = np.take_along_axis(cdf, below, 1)
    cdf_a = np.take_along_axis(cdf, above, 1)
    bin_b = np.take_along_axis(bins, below, 1)
    bin_a = np.take_along_axis(bins, above, 1)
    denom = cdf_a - cdf_b
    denom = np.where(denom < 1e-5, dtype(1.0), denom)
    t = (u - cdf_b) / denom
    samples = bin_b + t * (bin_a - bin_b)
    return samples, inds, cdf


# --------------------------------------------------------------------------
# A9  up_sample                                  models/renderer.py:132-176
# --------------------------------------------------------------------------


def up_sample_weights(rays_o, rays_d, z_vals, sdf, inv_s, dtype=np.float32):
    rays_o = np.asarray(rays_o, dtype)
    rays_d = np.asarray(rays_d, dtype)
    z_vals = np.asarray(z_vals, dtype)
    sdf = np.asarray(sdf, dtype)
    B, n = z_vals.shape
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[:, :, None]
    radius = np.sqrt((pts * pts).sum(-1))
    inside = (radius[:, :-1] < 1.0) | (radius[:, 1:] < 1.0)
    prev_sdf, next_sdf = sdf[:, :-1], sdf[:, 1:]
    prev_z, next_z = z_vals[:, :-1], z_vals[:, 1:]
    mid_sdf = (prev_sdf + next_sdf) * dtype(0.5)
    cos_val = (next_sdf - prev_sdf) / (next_z - prev_z + dtype(1e-5))
    prev_cos = np.concatenate([np.zeros((B, 1), dtype), cos_val[:, :-1]], -1)
    cos_val = np.minimum(prev_cos, cos_val)
    cos_val = np.clip(cos_val, -1e3, 0.0).astype(dtype) * inside
    dist = next_z - prev_z
    prev_esti = mid_sdf - cos_val * dist * dtype(0.5)
    next_esti = mid_sdf + cos_val * dist * dtype(0.5)
    prev_cdf = sigmoid(prev_esti * dtype(inv_s)).astype(dtype)
    next_cdf = sigmoid(next_esti * dtype(inv_s)).astype(dtype)
    alpha = (prev_cdf - next_cdf + dtype(1e-5)) / (prev_cdf + dtype(1e-5))
    T = np.cumprod(np.concatenate([np.ones((B, 1), dtype), 1.0 - alpha + dtype(1e-7)], -1), -1, dtype=dtype)[:, :-1]
    return (alpha * T).astype(dtype)


def up_sample(rays_o, rays_d, z_vals, sdf, n_importance, inv_s, dtype=np.float32):
    w = up_sample_weights(rays_o, rays_d, z_vals, sdf, inv_s, dtype)
    return sample_pdf_det(z_vals, w, n_importance, dtype)


# --------------------------------------------------------------------------
# A11  cat_z_vals                                models/renderer.py:178-192
# --------------------------------------------------------------------------


def cat_z_vals(z_vals, new_z, sdf=None, new_sdf=None):
    z = np.concatenate([z_vals, new_z], -1)
    idx = np.argsort(z, -1, kind="stable")
    z = np.take_along_axis(z, idx, -1)
    if sdf is None:
        return z, None
    s = np.concatenate([sdf, new_sdf], -1)
    return z, np.take_along_axis(s, idx, -1)


def hierarchical_sample(Ws, bs, rays_o, rays_d, near, far, t_rand, n_samples=64,
                        n_importance=64, up_sample_steps=4, dtype=np.float32):
    """The no_grad block of render_rnb* : models/renderer.py:829-880 (== 933-984).
    t_rand [B,1] is the value of `torch.rand([B,1]) - 0.5` (or None for perturb == 0)."""
    near = np.asarray(near, dtype)
    far = np.asarray(far, dtype)
    lin = np.linspace(0.0, 1.0, n_samples).astype(dtype)
    z = near + (far - near) * lin[None, :]
    if t_rand is not None:
        z = z + np.asarray(t_rand, dtype) * dtype(2.0 / n_samples)
    z = z.astype(dtype)
    B = z.shape[0]
    o = np.asarray(rays_o, dtype)
    d = np.asarray(rays_d, dtype)
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]
    sdf = sdf_only(Ws, bs, pts.reshape(-1, 3)).reshape(B, -1).astype(dtype)
    steps = []
    n_new = n_importance // up_sample_steps
    for i in range(up_sample_steps):
        new_z, inds, cdf = up_sample(o, d, z, sdf, n_new, 64 * 2 ** i, dtype)
        last = i + 1 == up_sample_steps
        if not last:
            p = o[:, None, :] + d[:, None, :] * new_z[:, :, None]
            new_sdf = sdf_only(Ws, bs, p.reshape(-1, 3)).reshape(B, -1).astype(dtype)
            z2, sdf2 = cat_z_vals(z, new_z, sdf, new_sdf)
        else:
            z2, sdf2 = cat_z_vals(z, new_z)
        steps.append(dict(z_in=z, sdf_in=sdf, new_z=new_z, inds=inds, cdf=cdf))
        z, sdf = z2, (sdf2 if sdf2 is not None else sdf)
    return z, steps


# --------------------------------------------------------------------------
# A12/A13  render_core_mvps + RNb shading       models/renderer.py:466-554, 904-930, 1008-1033
# --------------------------------------------------------------------------


def composite_forward(rays_o, rays_d, z_vals, sdf, grad, albedo, lights, inv_s,
                      cos_anneal_ratio, warmup, n_samples_cfg=64):
    """Everything of render_core_mvps after the networks, plus the shading sum.
    sdf [B,n], grad [B,n,3], albedo [B,n,3], lights [L,1,1,3] or [L,B,1,3]."""
    o = np.asarray(rays_o, F64)
    d = np.asarray(rays_d, F64)
    z = np.asarray(z_vals, F64)
    sdf = np.asarray(sdf, F64)
    grad = np.asarray(grad, F64)
    albedo = np.asarray(albedo, F64)
    B, n = z.shape
    sample_dist = 2.0 / n_samples_cfg
    dists = np.concatenate([z[:, 1:] - z[:, :-1], np.full((B, 1), sample_dist)], -1)
    mid = z + dists * 0.5
    pts = o[:, None, :] + d[:, None, :] * mid[:, :, None]
    true_cos = (d[:, None, :] * grad).sum(-1)
    r = cos_anneal_ratio
    iter_cos = -(np.maximum(-true_cos * 0.5 + 0.5, 0.0) * (1.0 - r) + np.maximum(-true_cos, 0.0) * r)
    est_next = sdf + iter_cos * dists * 0.5
    est_prev = sdf - iter_cos * dists * 0.5
    prev_cdf = sigmoid(est_prev * inv_s)
    next_cdf = sigmoid(est_next * inv_s)
    p = prev_cdf - next_cdf
    c = prev_cdf
    alpha_raw = (p + 1e-5) / (c + 1e-5)
    alpha = np.clip(alpha_raw, 0.0, 1.0)
    pts_norm = np.sqrt((pts * pts).sum(-1))
    inside = (pts_norm < 1.0).astype(F64)
    relax = (pts_norm < 1.2).astype(F64)
    T = np.cumprod(np.concatenate([np.ones((B, 1)), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    w = alpha * T
    gnorm = np.sqrt((grad * grad).sum(-1))
    eik_num = (relax * (gnorm - 1.0) ** 2).sum()
    eik_den = relax.sum() + 1e-5
    lights = np.broadcast_to(np.asarray(lights, F64), (lights.shape[0], B, 1, 3))
    shade = (grad[None] * lights).sum(-1)                      # [L,B,n]
    if warmup:
        shade = np.maximum(shade, 0.0)                         # models/renderer.py:912-913
    color = (albedo[None] * w[None, :, :, None] * shade[..., None]).sum(2)
    return dict(color_fine=color, weights=w, weight_sum=w.sum(-1, keepdims=True),
                weight_max=w.max(-1, keepdims=True), cdf_fine=c, inside_sphere=inside,
                gradient_error=eik_num / eik_den, s_val=np.full((B, 1), 1.0 / inv_s),
                dists=dists, mid_z_vals=mid, pts=pts, alpha=alpha, alpha_raw=alpha_raw, T=T,
                shade=shade, relax=relax, eik_den=eik_den, true_cos=true_cos,
                prev_cdf=prev_cdf, next_cdf=next_cdf, iter_cos=iter_cos, gnorm=gnorm)


def composite_backward(fw, rays_d, sdf, grad, albedo, lights, inv_s, cos_anneal_ratio, warmup,
                       d_color, d_weight_sum, d_eik):
    """Adjoint of composite_forward (SURVEY 8a' K5).  Returns d_sdf [B,n], d_grad [B,n,3],
    d_albedo [B,n,3], d_inv_s (scalar)."""
    d = np.asarray(rays_d, F64)
    sdf = np.asarray(sdf, F64)
    grad = np.asarray(grad, F64)
    albedo = np.asarray(albedo, F64)
    B, n = sdf.shape
    w, T, alpha = fw["weights"], fw["T"], fw["alpha"]
    shade = fw["shade"]
    lights = np.broadcast_to(np.asarray(lights, F64), (lights.shape[0], B, 1, 3))
    d_color = np.asarray(d_color, F64)                       # [L,B,3]
    # colour[l,b,c] = sum_i albedo[b,i,c] w[b,i] shade[l,b,i]
    d_albedo = np.einsum("lbc,bi,lbi->bic", d_color, w, shade)
    d_w = np.einsum("lbc,bic,lbi->bi", d_color, albedo, shade) + np.asarray(d_weight_sum, F64)
    d_shade = np.einsum("lbc,bic,bi->lbi", d_color, albedo, w)
    if warmup:
        d_shade = d_shade * (shade > 0)
    d_grad = (d_shade[..., None] * lights).sum(0)
    # w_i = alpha_i T_i ; T_i = prod_{j<i}(1-alpha_j+1e-7)
    ww = w * d_w
    suffix = np.cumsum(ww[:, ::-1], -1)[:, ::-1] - ww       # sum_{j>i} w_j dw_j
    d_alpha = T * d_w - suffix / (1.0 - alpha + 1e-7)
    ar = fw["alpha_raw"]
    d_alpha = d_alpha * ((ar >= 0.0) & (ar <= 1.0))          # clip passes gradient on the closed interval
    pc, nc = fw["prev_cdf"], fw["next_cdf"]
    # alpha_raw = (pc - nc + 1e-5)/(pc + 1e-5)
    d_pc = d_alpha * (1.0 / (pc + 1e-5) - (pc - nc + 1e-5) / (pc + 1e-5) ** 2)
    d_nc = -d_alpha / (pc + 1e-5)
    dists = fw["dists"]
    d_ep = d_pc * pc * (1.0 - pc)                            # d wrt (est_prev*inv_s)
    d_en = d_nc * nc * (1.0 - nc)
    est_prev = sdf - fw["iter_cos"] * dists * 0.5
    est_next = sdf + fw["iter_cos"] * dists * 0.5
    d_inv_s = (d_ep * est_prev + d_en * est_next).sum()
    d_est_prev = d_ep * inv_s
    d_est_next = d_en * inv_s
    d_sdf = d_est_prev + d_est_next
    d_iter_cos = (d_est_next - d_est_prev) * dists * 0.5
    tc = fw["true_cos"]
    r = cos_anneal_ratio
    d_tc = d_iter_cos * (0.5 * (1.0 - r) * ((-tc * 0.5 + 0.5) > 0) + r * ((-tc) > 0))
    d_grad = d_grad + d_tc[..., None] * d[:, None, :]
    # eikonal
    gn = fw["gnorm"]
    coef = d_eik * fw["relax"] * 2.0 * (gn - 1.0) / fw["eik_den"]
    safe = np.where(gn > 0, gn, 1.0)
    d_grad = d_grad + np.where(gn[..., None] > 0, coef[..., None] * grad / safe[..., None], 0.0)
    return d_sdf, d_grad, d_albedo, d_inv_s


# --------------------------------------------------------------------------
# A14  loss                                      exp_runner.py:241-256
# --------------------------------------------------------------------------


def rnb_loss(color_fine, weight_sum, gradient_error, true_rgb, mask, igr_weight=0.1, mask_weight=0.1):
    """Returns (loss, parts, cotangents) -- cotangents = d loss / d (color_fine, weight_sum, gradient_error)."""
    color_fine = np.asarray(color_fine, F64)
    true_rgb = np.asarray(true_rgb, F64)
    mask = np.asarray(mask, F64)
    L = color_fine.shape[0]
    mask_sum = mask.sum() + 1e-5
    err = (color_fine - true_rgb) * mask[None]
    color_loss = np.abs(err).sum() / (mask_sum * L)
    ws = np.clip(weight_sum, 1e-3, 1.0 - 1e-3)
    bce = -(mask * np.log(ws) + (1.0 - mask) * np.log(1.0 - ws)).mean()
    loss = color_loss + igr_weight * gradient_error + mask_weight * bce
    d_color = np.sign(err) * mask[None] / (mask_sum * L)
    inside = (np.asarray(weight_sum) >= 1e-3) & (np.asarray(weight_sum) <= 1.0 - 1e-3)
    d_ws = mask_weight * (-(mask / ws) + (1.0 - mask) / (1.0 - ws)) / mask.size * inside
    return loss, dict(color=color_loss, eik=gradient_error, mask=bce), (d_color, d_ws, igr_weight)


# --------------------------------------------------------------------------
# full step: render_rnb[_warmup] + loss + all parameter gradients
# --------------------------------------------------------------------------


def render_rnb(sdf_sd, color_sd, variance, rays_o, rays_d, near, far, lights, t_rand,
               cos_anneal_ratio=1.0, warmup=True, no_albedo=False, n_samples=64, n_importance=64,
               up_sample_steps=4, z_vals=None):
    """models/renderer.py:828-930 (warmup) / 932-1033.  Returns the output dict plus
    a cache for `train_step_grads`."""
    Ws, bs = sdf_effective(sdf_sd)
    if z_vals is None:
        z_vals, _ = hierarchical_sample(Ws, bs, rays_o, rays_d, near, far, t_rand, n_samples,
                                        n_importance, up_sample_steps)
    z = np.asarray(z_vals, F64)
    B, n = z.shape
    o = np.asarray(rays_o, F64)
    d = np.asarray(rays_d, F64)
    dists = np.concatenate([z[:, 1:] - z[:, :-1], np.full((B, 1), 2.0 / n_samples)], -1)
    mid = z + dists * 0.5
    pts = (o[:, None, :] + d[:, None, :] * mid[:, :, None]).reshape(-1, 3)
    out, g = sdf_gradient(Ws, bs, pts, keep=True)[:2]
    sdf = out[:, 0].reshape(B, n)
    feat = out[:, 1:]
    cWs = [weight_norm_fold(color_sd[f"lin{l}.weight_g"], color_sd[f"lin{l}.weight_v"]) for l in range(3)]
    cbs = [np.asarray(color_sd[f"lin{l}.bias"], F64) for l in range(3)]
    albedo_net = color_forward(cWs, cbs, pts, g, feat).reshape(B, n, 3)
    albedo = np.ones_like(albedo_net) if no_albedo else albedo_net
    inv_s = float(np.clip(np.exp(10.0 * float(variance)), 1e-6, 1e6))    # models/fields.py:323-325; renderer.py:503
    fw = composite_forward(o, d, z, sdf, g.reshape(B, n, 3), albedo, lights, inv_s, cos_anneal_ratio, warmup,
                           n_samples)
    ret = dict(color_fine=fw["color_fine"], s_val=fw["s_val"], cdf_fine=fw["cdf_fine"],
               weight_sum=fw["weight_sum"], weight_max=fw["weight_max"], gradients=g.reshape(B, n, 3),
               weights=fw["weights"], gradient_error=fw["gradient_error"], inside_sphere=fw["inside_sphere"])
    cache = dict(Ws=Ws, bs=bs, cWs=cWs, cbs=cbs, pts=pts, sdf=sdf, feat=feat, g=g, albedo=albedo, fw=fw,
                 inv_s=inv_s, z=z, B=B, n=n, lights=lights, warmup=warmup, no_albedo=no_albedo,
                 r=cos_anneal_ratio, d=d, variance=float(variance), sdf_sd=sdf_sd, color_sd=color_sd)
    return ret, cache


def train_step_grads(ret, cache, true_rgb, mask, igr_weight=0.1, mask_weight=0.1):
    """loss + gradient of every parameter, keyed like the reference state_dicts."""
    c = cache
    loss, parts, (d_color, d_ws, d_eik) = rnb_loss(ret["color_fine"], ret["weight_sum"], ret["gradient_error"],
                                                   true_rgb, mask, igr_weight, mask_weight)
    B, n = c["B"], c["n"]
    d_sdf, d_grad, d_albedo, d_inv_s = composite_backward(
        c["fw"], c["d"], c["sdf"], c["g"].reshape(B, n, 3), c["albedo"], c["lights"], c["inv_s"], c["r"],
        c["warmup"], d_color, d_ws, d_eik)
    grads = {}
    d_grad = d_grad.reshape(-1, 3)
    d_feat = np.zeros_like(c["feat"])
    if not c["no_albedo"]:
        cdW, cdb, d_n, d_f = color_backward(c["cWs"], c["cbs"], c["pts"], c["g"], c["feat"], d_albedo.reshape(-1, 3))
        d_grad = d_grad + d_n
        d_feat = d_f
        for l in range(3):
            dg, dv = weight_norm_vjp(c["color_sd"][f"lin{l}.weight_g"], c["color_sd"][f"lin{l}.weight_v"], cdW[l])
            grads[f"color.lin{l}.weight_g"] = dg
            grads[f"color.lin{l}.weight_v"] = dv
            grads[f"color.lin{l}.bias"] = cdb[l]
    ybar = np.concatenate([d_sdf.reshape(-1, 1), d_feat], 1)
    dWs, dbs = sdf_backward(c["Ws"], c["bs"], c["pts"], ybar, d_grad)
    for l in range(len(dWs)):
        dg, dv = weight_norm_vjp(c["sdf_sd"][f"lin{l}.weight_g"], c["sdf_sd"][f"lin{l}.weight_v"], dWs[l])
        grads[f"sdf.lin{l}.weight_g"] = dg
        grads[f"sdf.lin{l}.weight_v"] = dv
        grads[f"sdf.lin{l}.bias"] = dbs[l]
    # inv_s = clip(exp(10 v)) -> d v
    inside = 1e-6 <= np.exp(10.0 * c["variance"]) <= 1e6
    grads["variance"] = d_inv_s * 10.0 * c["inv_s"] * inside
    return loss, parts, grads, dict(d_sdf=d_sdf, d_grad=d_grad, d_albedo=d_albedo, d_feat=d_feat)


# --------------------------------------------------------------------------
# A16  extract_fields                            models/renderer.py:10-25, 1219-1224
# --------------------------------------------------------------------------


def grid_axes(bound_min, bound_max, resolution, dtype=np.float32):
    """torch.linspace(bmin, bmax, R) in float32 (models/renderer.py:12-14).
    torch computes start + i*step for the first half and end - (R-1-i)*step for the second."""
    axes = []
    for a in range(3):
        lo, hi = dtype(bound_min[a]), dtype(bound_max[a])
        step = dtype((hi - lo) / dtype(resolution - 1))
        i = np.arange(resolution)
        half = resolution // 2
        ax = np.where(i < half, lo + step * i.astype(dtype), hi - step * (resolution - 1 - i).astype(dtype)).astype(dtype)
        axes.append(ax)
    return axes


def extract_fields(Ws, bs, bound_min, bound_max, resolution, x_range=None, chunk=1 << 16):
    """u[x,y,z] = -sdf(X[x],Y[y],Z[z]) as float32, C-contiguous.  x_range=(x0,x1) evaluates one slab."""
    X, Y, Z = grid_axes(bound_min, bound_max, resolution)
    if x_range is not None:
        X = X[x_range[0]:x_range[1]]
    xx, yy, zz = np.meshgrid(X, Y, Z, indexing="ij")
    pts = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], -1)
    out = np.empty(pts.shape[0], np.float32)
    for s in range(0, pts.shape[0], chunk):
        out[s:s + chunk] = -sdf_only(Ws, bs, pts[s:s + chunk])[:, 0]
    return out.reshape(len(X), len(Y), len(Z))


# --------------------------------------------------------------------------
# A15  NeRF++ background                          models/fields.py:281-314; renderer.py:93-130
# --------------------------------------------------------------------------


def nerf_forward(sd, pts4, dirs, multires=10, multires_view=4, skips=(4,), D=8):
    e = embed(pts4, multires)
    ev = embed(dirs, multires_view)
    h = e
    for i in range(D):
        W = np.asarray(sd[f"pts_linears.{i}.weight"], F64)
        b = np.asarray(sd[f"pts_linears.{i}.bias"], F64)
        h = np.maximum(h @ W.T + b, 0.0)
        if i in skips:
            h = np.concatenate([e, h], -1)                     # models/fields.py:296-298
    alpha = h @ np.asarray(sd["alpha_linear.weight"], F64).T + np.asarray(sd["alpha_linear.bias"], F64)
    feat = h @ np.asarray(sd["feature_linear.weight"], F64).T + np.asarray(sd["feature_linear.bias"], F64)
    h = np.concatenate([feat, ev], -1)
    h = np.maximum(h @ np.asarray(sd["views_linears.0.weight"], F64).T + np.asarray(sd["views_linears.0.bias"], F64), 0.0)
    rgb = h @ np.asarray(sd["rgb_linear.weight"], F64).T + np.asarray(sd["rgb_linear.bias"], F64)
    return alpha, rgb


def render_core_outside(sd, rays_o, rays_d, z_vals, sample_dist):
    o = np.asarray(rays_o, F64)
    d = np.asarray(rays_d, F64)
    z = np.asarray(z_vals, F64)
    B, n = z.shape
    dists = np.concatenate([z[:, 1:] - z[:, :-1], np.full((B, 1), sample_dist)], -1)
    mid = z + dists * 0.5
    pts = o[:, None, :] + d[:, None, :] * mid[:, :, None]
    dis = np.clip(np.sqrt((pts * pts).sum(-1, keepdims=True)), 1.0, 1e10)
    pts4 = np.concatenate([pts / dis, 1.0 / dis], -1).reshape(-1, 4)
    dirs = np.broadcast_to(d[:, None, :], (B, n, 3)).reshape(-1, 3)
    density, rgb = nerf_forward(sd, pts4, dirs)
    color = sigmoid(rgb).reshape(B, n, 3)
    dens = density.reshape(B, n)
    sp = np.where(dens > 20.0, dens, np.log1p(np.exp(np.minimum(dens, 20.0))))
    alpha = 1.0 - np.exp(-sp * dists)
    return alpha, color


def render_plain_background(sdf_sd, color_sd, nerf_sd, variance, rays_o, rays_d, z_vals, z_feed,
                            cos_anneal_ratio=1.0, n_samples=64, background_rgb=None):
    """render() with n_outside > 0 after the sampling (models/renderer.py:609-648): render_core_outside on
    z_feed = sort(cat[z_vals, z_vals_outside]), then render_core with background_alpha /
    background_sampled_color (:255-260): inside_sphere blend of alpha and colour, the outside samples appended."""
    Ws, bs = sdf_effective(sdf_sd)
    z = np.asarray(z_vals, F64)
    B, n = z.shape
    o = np.asarray(rays_o, F64)
    d = np.asarray(rays_d, F64)
    sample_dist = 2.0 / n_samples
    bg_alpha, bg_color = render_core_outside(nerf_sd, o, d, z_feed, sample_dist)
    dists = np.concatenate([z[:, 1:] - z[:, :-1], np.full((B, 1), sample_dist)], -1)
    mid = z + dists * 0.5
    pts = (o[:, None, :] + d[:, None, :] * mid[:, :, None]).reshape(-1, 3)
    out, g = sdf_gradient(Ws, bs, pts, keep=True)[:2]
    sdf = out[:, 0].reshape(B, n)
    cWs = [weight_norm_fold(color_sd[f"lin{l}.weight_g"], color_sd[f"lin{l}.weight_v"]) for l in range(3)]
    cbs = [np.asarray(color_sd[f"lin{l}.bias"], F64) for l in range(3)]
    color = color_forward(cWs, cbs, pts, g, out[:, 1:]).reshape(B, n, 3)
    inv_s = float(np.clip(np.exp(10.0 * float(variance)), 1e-6, 1e6))
    ones = np.zeros((1, 1, 1, 3))
    fw = composite_forward(o, d, z, sdf, g.reshape(B, n, 3), color, ones, inv_s, cos_anneal_ratio, False, n_samples)
    inside = fw["inside_sphere"]
    alpha = fw["alpha"] * inside + bg_alpha[:, :n] * (1.0 - inside)
    alpha = np.concatenate([alpha, bg_alpha[:, n:]], -1)
    col = color * inside[:, :, None] + bg_color[:, :n] * (1.0 - inside)[:, :, None]
    col = np.concatenate([col, bg_color[:, n:]], 1)
    T = np.cumprod(np.concatenate([np.ones((B, 1)), 1.0 - alpha + 1e-7], -1), -1)[:, :-1]
    w = alpha * T
    wsum = w.sum(-1, keepdims=True)
    c = (col * w[:, :, None]).sum(1)
    if background_rgb is not None:
        c = c + np.asarray(background_rgb, F64) * (1.0 - wsum)
    return dict(color_fine=c, weights=w, weight_sum=wsum, weight_max=w.max(-1, keepdims=True),
                cdf_fine=fw["cdf_fine"], inside_sphere=inside, gradient_error=fw["gradient_error"],
                gradients=g.reshape(B, n, 3), s_val=fw["s_val"], bg_alpha=bg_alpha, bg_color=bg_color)


# --------------------------------------------------------------------------
# 8f-2  step epilogue: Adam + learning-rate schedule   exp_runner.py:115, 263, 320-332
# --------------------------------------------------------------------------

def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """One torch.optim.Adam update (amsgrad off, weight_decay 0: what exp_runner.py:115 constructs), `step` = 1-based
    count of this update.  Returns the new (p, m, v); float64."""
    p, g, m, v = (np.asarray(a, F64) for a in (p, g, m, v))
    m = m + (g - m) * (1.0 - beta1)
    v = v * beta2 + (1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = np.sqrt(v) / np.sqrt(bc2) + eps
    return p - (lr / bc1) * (m / denom), m, v


def learning_rate(iter_step, base_lr=5e-4, alpha=0.05, warm_up_end=5000, end_iter=300000):
    """exp_runner.py:320-332 (defaults: confs/wmask_rnb.conf:21-23, 28)."""
    if iter_step < warm_up_end:
        f = iter_step / warm_up_end
    else:
        progress = (iter_step - warm_up_end) / (end_iter - warm_up_end)
        f = (np.cos(np.pi * progress) + 1.0) * 0.5 * (1 - alpha) + alpha
    return base_lr * f
