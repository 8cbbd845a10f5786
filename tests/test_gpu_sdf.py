"""GPU parity of the SDF kernels (through the C ABI) against the numpy oracle and the golden fixtures.
Tolerance: SDF / gradient <= 1e-3 relative in the norm (north_star); fp16 operands, fp32 accumulate."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2, check
from gpu_common import build_nets, np_state
from oracle import rnb_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-3


def packed_for(sdf):
    from rnb_b200 import kernels as K
    eff = sdf.effective_weights()
    return K.SdfPacked("cuda").pack([w for w, _ in eff], [b for _, b in eff])


@pytest.mark.parametrize("perturb", [False, True])
@pytest.mark.parametrize("n", [96, 1000, 5000])
def test_sdf_fwd_and_grad(perturb, n):
    from rnb_b200 import kernels as K
    _, sdf, _, _ = build_nets(perturb)
    pk = packed_for(sdf)
    g = torch.Generator().manual_seed(n)
    x = ((torch.rand(n, 3, generator=g) - 0.5) * 2.4).cuda()
    Ws, bs = O.sdf_effective(np_state(sdf))
    ref_out, ref_grad = O.sdf_gradient(Ws, bs, x.cpu().numpy(), keep=True)[:2]
    out = K.sdf_fwd(pk, K.points_explicit(x)).cpu().numpy()
    assert rel_l2(out, ref_out[:, 0]) < TOL
    s, gr, full, st = K.sdf_fwd_grad(pk, K.points_explicit(x), want_full=True)
    torch.cuda.synchronize()
    assert rel_l2(s.cpu().numpy(), ref_out[:, 0]) < TOL
    assert rel_l2(gr.cpu().numpy(), ref_grad) < TOL
    assert rel_l2(full.cpu().numpy(), ref_out) < TOL
    feat = K.stream_to_rowmajor(st.feat, n, 256).float().cpu().numpy()
    check("sdf: feature stream (fp16 storage) vs oracle", rel_l2(feat, ref_out[:, 1:]), 1e-3)


@pytest.mark.parametrize("perturb", [False, True])
def test_sdf_golden(perturb):
    from rnb_b200 import kernels as K
    gld = load_golden("sdf_perturbed" if perturb else "sdf_init")
    _, sdf, _, _ = build_nets(perturb)
    pk = packed_for(sdf)
    x = torch.from_numpy(gld["x"]).cuda()
    s, gr, full, _ = K.sdf_fwd_grad(pk, K.points_explicit(x), want_full=True)
    assert rel_l2(full.cpu().numpy(), gld["out"]) < TOL
    assert rel_l2(gr.cpu().numpy(), gld["grad"]) < TOL


@pytest.mark.parametrize("perturb", [False, True])
@pytest.mark.parametrize("n", [200, 4000])
def test_sdf_backward(perturb, n):
    """Double-backward: parameter gradients cos >= 0.999 and rel-L2 <= 1e-2 (north_star)."""
    from conftest import cosine
    from rnb_b200 import kernels as K
    _, sdf, _, _ = build_nets(perturb)
    pk = packed_for(sdf)
    g = torch.Generator().manual_seed(100 + n)
    x = ((torch.rand(n, 3, generator=g) - 0.5) * 2.0).cuda()
    # cotangent magnitudes like a real RNb loss (SURVEY 8a': d_sdf ~1e-6..1e-3, d_feat ~1e-10..1e-6)
    d_sdf = (torch.randn(n, generator=g) * 1e-4 * torch.rand(n, generator=g) ** 4).cuda()
    d_grad = (torch.randn(n, 3, generator=g) * 3e-5).cuda()
    d_feat = (torch.randn(n, 256, generator=g) * 1e-7 * (torch.rand(n, 256, generator=g) > 0.33)).cuda()
    pts = K.points_explicit(x)
    _, _, _, st = K.sdf_fwd_grad(pk, pts)
    dWs, dbs, _ = K.sdf_bwd(pk, pts, st, d_sdf, d_grad, d_feat)
    torch.cuda.synchronize()
    Ws, bs = O.sdf_effective(np_state(sdf))
    ybar = np.concatenate([d_sdf.cpu().numpy()[:, None], d_feat.cpu().numpy()], 1)
    rW, rb = O.sdf_backward(Ws, bs, x.cpu().numpy(), ybar, d_grad.cpu().numpy())
    for l in range(9):
        for nm, got, ref in ((f"dW{l}", dWs[l], rW[l]), (f"db{l}", dbs[l], rb[l])):
            got = got.cpu().numpy()
            assert np.isfinite(got).all(), nm
            assert cosine(got, ref) > 0.999, (nm, cosine(got, ref))
            check(f"sdf backward: {nm} vs oracle", rel_l2(got, ref), 1e-2)


@pytest.mark.parametrize("perturb", [False, True])
def test_albedo_fwd_bwd(perturb):
    """Albedo net on the SDF kernel's own feature stream, against the oracle (colour tolerance 1e-3; grads 1e-2)."""
    from conftest import cosine
    from rnb_b200 import kernels as K, albedo as A
    n = 3000
    _, sdf, _, col = build_nets(perturb)
    pk = packed_for(sdf)
    g = torch.Generator().manual_seed(7)
    x = ((torch.rand(n, 3, generator=g) - 0.5) * 2.0).cuda()
    pts = K.points_explicit(x)
    _, grad, full, st = K.sdf_fwd_grad(pk, pts, want_full=True)
    flat = []
    for W, b in col.effective_weights():
        flat += [W, b]
    ctx = A.forward(flat, pts, grad, st)
    torch.cuda.synchronize()
    cs = np_state(col)
    cWs = [O.weight_norm_fold(cs[f"lin{l}.weight_g"], cs[f"lin{l}.weight_v"]) for l in range(3)]
    cbs = [cs[f"lin{l}.bias"] for l in range(3)]
    feat16 = K.stream_to_rowmajor(st.feat, n, 256).float().cpu().numpy()     # the fp16 features the kernel consumed
    ref = O.color_forward(cWs, cbs, x.cpu().numpy(), grad.cpu().numpy(), feat16)
    assert rel_l2(ctx.albedo.cpu().numpy(), ref) < 1e-3
    ref_full = O.color_forward(cWs, cbs, x.cpu().numpy(), grad.cpu().numpy(), full[:, 1:].cpu().numpy())
    assert rel_l2(ctx.albedo.cpu().numpy(), ref_full) < 1e-3
    d_alb = (torch.randn(n, 3, generator=g) * 1e-4 * torch.rand(n, 1, generator=g) ** 3).cuda()
    d_normal, d_feat, grads = A.backward(ctx, d_alb, want_fp32=True)
    torch.cuda.synchronize()
    # the training path hands d_feat to the SDF backward as an fp16 stream in the albedo backward's power-of-two
    # cotangent scale (no fp32 [n,256] round trip): same values, and the recorded maximum is the stream's maximum
    _, (d_feat16, meta), _ = A.backward(ctx, d_alb)
    import math
    mant, expo = math.frexp(float(meta[0]))
    scale_alb = 2.0 ** (8 - expo)
    assert abs(float(meta[0]) - float(d_alb.abs().max())) < 1e-12
    stored = K.stream_to_rowmajor(d_feat16, n, 256).float()
    assert abs(float(stored.abs().max()) - float(meta[1])) <= 2e-3 * float(meta[1])
    assert rel_l2((stored / scale_alb).cpu().numpy(), d_feat.cpu().numpy()) < 2e-3
    # (i) backward arithmetic, with the kernel's own ReLU masks (see oracle.color_backward)
    masks = [K.stream_to_rowmajor(ctx.st_h0, n, 256).float().cpu().numpy() > 0,
             K.stream_to_rowmajor(ctx.st_h1, n, 256).float().cpu().numpy() > 0]
    mW, mb, rn, rf = O.color_backward(cWs, cbs, x.cpu().numpy(), grad.cpu().numpy(), feat16, d_alb.cpu().numpy(),
                                      relu_masks=masks)
    assert rel_l2(d_normal.cpu().numpy(), rn) < 1e-2, rel_l2(d_normal.cpu().numpy(), rn)
    assert rel_l2(d_feat.cpu().numpy(), rf) < 1e-2, rel_l2(d_feat.cpu().numpy(), rf)
    # (ii) against the exact oracle: per-point cotangents carry the mask-flip noise (a few 1e-4 of the units),
    # parameter gradients (sums over points) must still meet the gradient criterion
    rW, rb, rn2, rf2 = O.color_backward(cWs, cbs, x.cpu().numpy(), grad.cpu().numpy(), feat16, d_alb.cpu().numpy())
    assert cosine(d_normal.cpu().numpy(), rn2) > 0.999 and cosine(d_feat.cpu().numpy(), rf2) > 0.999
    for l in range(3):
        for nm, got, want, exact in ((f"dW{l}", grads[2 * l], mW[l], rW[l]), (f"db{l}", grads[2 * l + 1], mb[l], rb[l])):
            got = got.cpu().numpy()
            assert rel_l2(got, want) < 1e-2, (nm, rel_l2(got, want))            # arithmetic, same masks
            assert cosine(got, exact) > 0.999, (nm, cosine(got, exact))         # exact oracle
            check(f"albedo backward: {nm} vs exact oracle (3000 points, sparse cotangents)", rel_l2(got, exact), 2e-2)


def test_fused_weight_norm_matches_oracle_and_torch():
    """rnb_weight_norm_fold / _vjp (all layers in one launch each) vs the oracle (models/fields.py:72-74 restated) and vs
    torch._weight_norm + autograd, on the shapes of the SDF and albedo networks."""
    from oracle import rnb_oracle as O
    from rnb_b200 import wnorm
    g0 = torch.Generator().manual_seed(3)
    shapes = [(256, 39), (256, 256), (217, 256), (257, 256), (256, 310), (3, 256)]
    vs = [torch.randn(s, generator=g0).cuda().requires_grad_(True) for s in shapes]
    gs = [(torch.rand(s[0], 1, generator=g0) + 0.5).cuda().requires_grad_(True) for s in shapes]
    dWs = [torch.randn(s, generator=g0).cuda() for s in shapes]
    Ws = wnorm.fold_all(vs, gs)
    torch.autograd.backward(Ws[:5], dWs[:5])                     # the last layer receives no gradient (must stay None)
    vt = [v.detach().clone().requires_grad_(True) for v in vs]
    gt = [g.detach().clone().requires_grad_(True) for g in gs]
    Wt = [torch._weight_norm(v, g, 0) for v, g in zip(vt, gt)]
    torch.autograd.backward(Wt[:5], dWs[:5])
    for i in range(6):
        Wo = O.weight_norm_fold(gs[i].detach().cpu().numpy(), vs[i].detach().cpu().numpy())
        assert rel_l2(Ws[i].detach().cpu().numpy(), Wo) < 1e-6
        assert torch.allclose(Ws[i], Wt[i], rtol=1e-5, atol=1e-7)
        if i == 5:
            assert vs[i].grad is None and gs[i].grad is None      # no cotangent: no gradient, not a zero tensor
            continue
        dgo, dvo = O.weight_norm_vjp(gs[i].detach().cpu().numpy(), vs[i].detach().cpu().numpy(), dWs[i].cpu().numpy())
        assert rel_l2(vs[i].grad.cpu().numpy(), dvo) < 1e-5 and rel_l2(gs[i].grad.cpu().numpy(), dgo) < 1e-5
        assert gs[i].grad.shape == gs[i].shape
        assert rel_l2(vs[i].grad.cpu().numpy(), vt[i].grad.cpu().numpy()) < 1e-5
