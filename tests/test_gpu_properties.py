"""Edge cases and size-independent properties of the public path on the GPU (BASELINE.json full sizes included):
ragged / tiny / empty batches, determinism, exact linearity of the backward in the cotangents, physical bounds of the
compositing outputs, and a short optimisation run."""
import numpy as np
import pytest
import torch

from gpu_common import build_nets
from rnb_b200 import synth
from test_gpu_e2e import loss_fn, make_renderer

pytestmark = pytest.mark.gpu


def batch(B, seed=1):
    return {k: v.cuda() for k, v in synth.make_batch(B, 3, True, seed).items()}


def render(renderer, b, **kw):
    return renderer.render_rnb_warmup(b["rays_o"], b["rays_d"], b["near"], b["far"], b["lights_dir"], cos_anneal_ratio=1.0, **kw)


@pytest.mark.parametrize("B", [1, 5, 131])
def test_tiny_and_ragged_batches(B):
    """ray counts that fill neither a 4-ray compositing block nor a 128-point MLP tile of the coarse passes"""
    renderer, sdf, var, col = make_renderer(True)
    b = batch(B)
    out = render(renderer, b)
    assert tuple(out["color_fine"].shape) == (3, B, 3) and tuple(out["weights"].shape) == (B, 128)
    assert tuple(out["gradients"].shape) == (B, 128, 3) and tuple(out["weight_sum"].shape) == (B, 1)
    for k in ("color_fine", "weights", "weight_sum", "gradients", "cdf_fine", "gradient_error"):
        assert torch.isfinite(out[k]).all(), k
    loss_fn(out, b["true_rgb"], b["mask"], 0.1).backward()
    for m in (sdf, var, col):
        for n, p in m.named_parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all(), n


def test_empty_point_sets():
    from rnb_b200 import kernels as K, ops
    _, sdf, _, _ = build_nets(True)
    out = sdf.sdf(torch.empty(0, 3, device="cuda"))
    assert tuple(out.shape) == (0, 1)
    pk = ops.packed_sdf_nograd(sdf)
    s, g, full, _ = K.sdf_fwd_grad(pk, K.points_explicit(torch.empty(0, 3, device="cuda")), want_full=True)
    assert s.numel() == 0 and tuple(g.shape) == (0, 3) and tuple(full.shape) == (0, 257)


def test_full_size_bounds_determinism_linearity():
    """BASELINE.json configs[3] size (8192 rays, 1 M fine points)"""
    renderer, sdf, var, col = make_renderer(True)
    b = batch(8192)
    params = [p for m in (sdf, var, col) for p in m.parameters()]

    def run(scale):
        for p in params:
            p.grad = None
        torch.manual_seed(3)
        out = render(renderer, b)
        (loss_fn(out, b["true_rgb"], b["mask"], 0.1) * scale).backward()
        return out, [p.grad.clone() for p in params]

    out, g1 = run(1.0)
    w = out["weights"]
    assert (w >= 0).all() and (out["weight_sum"] <= 1.0 + 1e-4).all() and (out["weight_max"] <= out["weight_sum"] + 1e-6).all()
    assert ((out["cdf_fine"] >= 0) & (out["cdf_fine"] <= 1)).all()
    assert set(np.unique(out["inside_sphere"].cpu().numpy())) <= {0.0, 1.0}
    assert torch.allclose(w.sum(-1, keepdim=True), out["weight_sum"], atol=1e-5)
    assert float(out["gradient_error"]) >= 0
    # determinism: fixed split-K order and partial sums -> bitwise identical outputs and gradients
    out2, g2 = run(1.0)
    assert torch.equal(out["color_fine"], out2["color_fine"]) and torch.equal(out["gradients"], out2["gradients"])
    for a, c in zip(g1, g2):
        assert torch.equal(a, c)
    # linearity: the backward is linear in the cotangents, and its fp16 operands carry a power-of-two scale derived from
    # max|cotangent| -- scaling the loss by 4 scales every kernel-side gradient by exactly 4
    _, g4 = run(4.0)
    for a, c in zip(g1, g4):
        assert torch.allclose(4.0 * a, c, rtol=1e-5, atol=1e-12)


def test_short_optimisation_run_reduces_loss():
    renderer, sdf, var, col = make_renderer(False)
    params = [p for m in (sdf, var, col) for p in m.parameters()]
    opt = torch.optim.Adam(params, lr=5e-4)
    b = batch(2048, seed=5)
    losses = []
    for it in range(12):
        opt.zero_grad()
        torch.manual_seed(it)
        loss = loss_fn(render(renderer, b), b["true_rgb"], b["mask"], 0.1)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert np.isfinite(losses).all()
    assert losses[-1] < losses[0], losses


def test_no_grad_fast_path_matches_training_forward():
    """validate_image renders under torch.no_grad(): same outputs as the autograd path, without the backward's streams"""
    renderer, sdf, var, col = make_renderer(True)
    b = batch(517)
    for no_albedo in (False, True):
        torch.manual_seed(5)
        ref = render(renderer, b, no_albedo=no_albedo)
        with torch.no_grad():
            torch.manual_seed(5)
            out = render(renderer, b, no_albedo=no_albedo)
        for k in ("color_fine", "weights", "weight_sum", "weight_max", "gradients", "cdf_fine", "inside_sphere", "gradient_error"):
            assert torch.equal(out[k], ref[k].detach()), k
            assert not out[k].requires_grad


def test_graphed_step_matches_eager(monkeypatch):
    """the whole step as one CUDA graph (launch-bound 512-ray regime): same loss and gradients as the eager step, also
    after optimiser steps in between (the weight packing is part of the graph)"""
    # capture folds the weight norm with the fused node; give the eager step the same node so the comparison stays bit-exact
    # (fused vs per-layer torch fold is compared in test_fused_weight_norm_step_matches_torch_weight_norm)
    monkeypatch.setenv("RNB_FUSED_WN", "1")
    from rnb_b200.graph_step import GraphedTrainStep
    from rnb_b200.parallel import FlatGradAllReducer
    renderer, sdf, var, col = make_renderer(True)
    renderer.perturb = 0.0                       # no jitter: eager and graphed steps must agree bit for bit
    params = [p for m in (sdf, var, col) for p in m.parameters()]
    red = FlatGradAllReducer(params)
    opt = torch.optim.Adam(params, lr=1e-3)
    lf = lambda out, rgb, mask: loss_fn(out, rgb, mask, 0.1)
    gs = GraphedTrainStep(renderer, params, lf, batch(512, seed=2), warmup=True, reducer=red)
    for seed in (3, 4, 5):
        b = batch(512, seed=seed)
        loss_g = gs(b).clone()
        grads_g = red.flat.clone()
        red.zero()
        loss_e = lf(render(renderer, b), b["true_rgb"], b["mask"])
        loss_e.backward()
        assert torch.equal(loss_g, loss_e.detach())
        assert torch.equal(grads_g, red.collect())
        del loss_e                               # keep no eager autograd graph alive across captures / replays
        opt.step()                               # parameters change in place: the next replay must see them
    # with jitter every replay draws new random numbers
    renderer.perturb = 1.0
    gs2 = GraphedTrainStep(renderer, params, lf, batch(512, seed=2), warmup=True, reducer=red)
    b = batch(512, seed=6)
    l1 = gs2(b).clone()
    l2 = gs2(b).clone()
    assert torch.isfinite(l1) and torch.isfinite(l2) and not torch.equal(l1, l2)


def test_fused_weight_norm_step_matches_torch_weight_norm(monkeypatch):
    """the whole train step with the fused weight-norm node (used under CUDA-graph capture) gives the gradients of the
    per-layer torch path, to the bar SURVEY 8c sets for parameter gradients (cosine >= 0.999, rel-L2 <= 1e-2).
    The two folds round W differently in the last fp32 bit; with beta = 100 softplus and fp16 operands the step amplifies
    that: moving every weight_g by ONE fp32 ulp under the torch fold changes the gradients by the same amount as switching
    folds (profiles/_wn_diag.py, 4096 rays: <= 4.7e-3 rel-L2 either way; at 256 rays 1.2e-2 vs 0.8e-2).  So the comparison
    runs at 4096 rays and is not bit for bit; the node itself is pinned at 1e-5 in
    test_gpu_sdf.py::test_fused_weight_norm_matches_oracle_and_torch."""
    renderer, sdf, var, col = make_renderer(True)
    renderer.perturb = 0.0
    params = [p for m in (sdf, var, col) for p in m.parameters()]
    b = batch(4096, seed=3)
    res = []
    for flag in ("0", "1"):
        monkeypatch.setenv("RNB_FUSED_WN", flag)
        for p in params:
            p.grad = None
        torch.manual_seed(11)
        loss = loss_fn(render(renderer, b), b["true_rgb"], b["mask"], 0.1)
        loss.backward()
        res.append((float(loss.detach()), [p.grad.clone() for p in params]))
    assert abs(res[0][0] - res[1][0]) < 1e-4 * abs(res[0][0])
    for g0, g1 in zip(res[0][1], res[1][1]):
        assert g0.shape == g1.shape
        n0, n1 = float(g0.norm()), float(g1.norm())
        assert float((g0 - g1).norm()) <= 1e-2 * n0 + 1e-12
        assert float((g0 * g1).sum()) >= 0.999 * n0 * n1
