"""End-to-end parity of NeuSRenderer.render_rnb[_warmup] (+ loss + backward) and extract_fields on the GPU:
against the golden fixtures written by the reference and against the oracle.
Tolerances (north_star): outputs / loss <= 1e-3 relative; parameter gradients cos >= 0.999, rel-L2 <= 1e-2."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_l2, cosine, check
from gpu_common import build_nets, np_state
from oracle import rnb_oracle as O
from rnb_b200 import synth

TOL_OWN_SAMPLING = 1e-3      # north_star; measured worst 7.3e-4 (profiles/r02_parity_margins.md)

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def make_renderer(perturb):
    from models.renderer import NeuSRenderer
    nerf, sdf, var, col = build_nets(perturb)
    r = NeuSRenderer(nerf, sdf, var, col, **synth.WMASK_CONF["neus_renderer"])
    r.color_depth = 3
    return r, sdf, var, col


def loss_fn(out, true_rgb, mask, mask_weight):
    # reference exp_runner.py:241-256
    mask_sum = mask.sum() + 1e-5
    err = ((out["color_fine"] - true_rgb) * mask[None, :, :]).reshape(-1, 3)
    color_loss = F.l1_loss(err, torch.zeros_like(err), reduction="sum") / (mask_sum * true_rgb.shape[0])
    mask_loss = F.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), mask)
    return color_loss + out["gradient_error"] * 0.1 + mask_loss * mask_weight


class FixedRand:
    """Inject the jitter the fixture was generated with (CPU and CUDA generators differ)."""

    def __init__(self, t):
        self.t, self.orig = t, torch.rand

    def __enter__(self):
        torch.rand = lambda *a, **k: (self.t + 0.5).to(k.get("device", "cpu"))

    def __exit__(self, *a):
        torch.rand = self.orig


CASES_NOALBEDO = ["warmup_noalbedo", "post_noalbedo"]
CASES_ALBEDO = ["warmup_albedo", "post_albedo", "womask_anneal", "init_warmup_albedo"]


@pytest.mark.parametrize("case", CASES_NOALBEDO + CASES_ALBEDO)
def test_render_rnb_golden(case):
    g = load_golden("render_" + case)
    renderer, sdf, var, col = make_renderer(not case.startswith("init"))
    warm, no_albedo = bool(g["warmup"]), bool(g["no_albedo"])
    fn = renderer.render_rnb_warmup if warm else renderer.render_rnb
    args = (cu(g["rays_o"]), cu(g["rays_d"]), cu(g["near"]), cu(g["far"]), cu(g["lights_dir"]))
    # (a) the public call with its own hierarchical sampling: ray-integrated outputs and the loss
    with FixedRand(torch.from_numpy(g["t_rand"])):
        out = fn(*args, cos_anneal_ratio=float(g["r"]), no_albedo=no_albedo)
    # The importance samples come from the library's own sampling pass (fp16-operand SDF), so they sit ~1e-4 away from
    # the reference's; the ray integrals still hold the north_star 1e-3 (measured worst 7.3e-4 over the six fixtures,
    # profiles/r02_parity_margins.md; tests/test_oracle_golden.py::test_sampling_pass_sensitivity bounds the effect of
    # such a shift in the float64 oracle).
    for k in ("color_fine", "weight_sum", "s_val"):
        assert tuple(out[k].shape) == g["out_" + k].shape, k
        check(f"e2e[{case}] public call, own sampling: {k} vs reference", rel_l2(out[k].detach().cpu().numpy(), g["out_" + k]), TOL_OWN_SAMPLING)
    # the eikonal term is a mean over the sample positions themselves; rays that miss the surface have flat weight
    # profiles, so a 1e-4 SDF difference in the sampling pass moves their importance samples a lot and the 16-ray mean by
    # up to ~5e-3 (the oracle shows the same sensitivity: test_sampling_pass_sensitivity) -- checked at 1e-3 in (b)
    check(f"e2e[{case}] public call, own sampling: eikonal term vs reference", abs(float(out["gradient_error"]) / float(g["out_gradient_error"]) - 1), 1e-2)
    loss_a = loss_fn(out, cu(g["true_rgb"]), cu(g["mask_used"]), float(g["mask_weight"]))
    assert abs(float(loss_a) / float(g["loss"]) - 1) < 1e-3
    # (b) identical inputs for the fine pass (the reference's own z_vals): per-sample outputs are comparable only
    # then -- a 1e-4 shift of a sample changes its section length, hence its individual weight, by several percent
    out = renderer._render_rnb(warm, *args, -1, None, float(g["r"]), no_albedo, _z_vals=cu(g["z_vals"]))
    for k in ("color_fine", "weight_sum", "weight_max", "weights", "cdf_fine", "gradients", "s_val"):
        assert tuple(out[k].shape) == g["out_" + k].shape, k
        # north_star tolerance 1e-3 for SDF / colour / normals / loss.  The individual sample weights are not in that
        # list and cannot be: alpha differences two sigmoids of inv_s * sdf, so an SDF that is right to 3e-4
        # relative moves a single weight by inv_s/10 times that; the ray sums (colour, weight_sum) stay at 1e-3.
        tol = 1e-2 if k in ("weights", "weight_max") else 1e-3
        check(f"e2e[{case}] reference z_vals: {k} vs reference", rel_l2(out[k].detach().cpu().numpy(), g["out_" + k]), tol)
    check(f"e2e[{case}] reference z_vals: eikonal term", abs(float(out["gradient_error"]) / float(g["out_gradient_error"]) - 1), 1e-3)
    mism = (out["inside_sphere"].cpu().numpy() != g["out_inside_sphere"]).mean()
    assert mism < 1e-3
    loss = loss_fn(out, cu(g["true_rgb"]), cu(g["mask_used"]), float(g["mask_weight"]))
    assert abs(float(loss) / float(g["loss"]) - 1) < 1e-3
    loss.backward()
    st = int(g["stride"])
    n = 0
    all_got, all_ref = [], []
    for tag, mod in (("sdf", sdf), ("color", col), ("var", var)):
        for pname, p in sorted(mod.named_parameters()):
            key = f"g_{tag}.{pname}"
            if key not in g:
                if tag == "color" and no_albedo:
                    assert p.grad is None           # reference: colour net receives no gradient with --no_albedo
                continue
            assert p.grad is not None, key
            got = p.grad.detach().cpu().numpy().reshape(-1)
            ref = g[key]
            if float(g[f"n_{tag}.{pname}"]) < 1e-12:
                continue
            assert np.isfinite(got).all(), key
            all_got.append(got[::st])
            all_ref.append(ref)
            # One-degree-of-freedom gradients (a tensor whose energy sits in one element: the sdf row of lin8.bias /
            # lin8.weight_g, the variance) are sums over all points with heavy cancellation -- d loss / d (scale of
            # the sdf row) is the small residual of two large sums, and rays that miss contribute through sigmoid
            # tails exp(-inv_s*sdf) where a 1e-4 absolute SDF error (3e-4 relative, inside the SDF tolerance) is
            # a 1 % error.  They are checked at 1e-1 on this fixture and contribute to the whole-vector check.
            # The albedo net is a ReLU MLP: the fp16 forward flips the sign of the few pre-activations within its
            # rounding error of zero (a few 1e-4 of the units) and each flip changes that point's cotangent by
            # 100 %.  Summed over the 2048 points of this fixture that leaves 1-2 % on its first-layer gradients
            # (3e-2 here); test_render_rnb_oracle_large checks 1e-2 at a realistic point count.
            single = float(np.abs(ref).max()) ** 2 > 0.99 * float((ref.astype(np.float64) ** 2).sum())
            # Per-tensor bounds on this 16-ray fixture, set from the measured errors (profiles/r02_parity_margins.md): SDF
            # tensors 1.25e-2 (worst 8.5e-3: weight_g gradients are row-wise projections of dW onto v with cancellation and
            # few contributing points), albedo net 2e-2 (worst 1.5e-2, relu flips), one-element tensors 8e-2 (worst 5.8e-2).
            # The north_star bound 1e-2 is asserted on the whole parameter vector below (measured 3.2e-3) and per tensor at
            # 192 and 512 rays (test_render_rnb_oracle_large, tests/test_gpu_large.py).
            tol = 8e-2 if single else (2e-2 if tag == "color" else 1.25e-2)
            assert cosine(got[::st], ref) > 0.999, (key, cosine(got[::st], ref))
            check(f"e2e[{case}] gradient {key} (16 rays{', one-element tensor' if single else ''})", rel_l2(got[::st], ref), tol)
            assert abs(np.linalg.norm(got) / float(g[f"n_{tag}.{pname}"]) - 1) < tol, key
            n += 1
    assert n >= 16
    all_got, all_ref = np.concatenate(all_got), np.concatenate(all_ref)
    assert cosine(all_got, all_ref) > 0.999
    check(f"e2e[{case}] whole parameter-gradient vector (16 rays)", rel_l2(all_got, all_ref), 1e-2)


@pytest.mark.parametrize("warm,no_albedo", [(True, False), (False, True)])
def test_render_rnb_oracle_large(warm, no_albedo):
    """192 rays (24576 fine points): every parameter gradient against the float64 oracle at the north_star
    tolerance (cos >= 0.999, rel-L2 <= 1e-2), colour / loss at 1e-3; same sample depths on both sides."""
    B = 192
    renderer, sdf, var, col = make_renderer(True)
    b = {k: v.cuda() for k, v in synth.make_batch(B, 3, warm, 11).items()}
    with torch.no_grad():
        z_vals, _ = renderer._sample(b["rays_o"], b["rays_d"], b["near"], b["far"], -1)
    out = renderer._render_rnb(warm, b["rays_o"], b["rays_d"], b["near"], b["far"], b["lights_dir"], -1, None, 1.0,
                               no_albedo, _z_vals=z_vals)
    loss = loss_fn(out, b["true_rgb"], b["mask"], 0.1)
    loss.backward()
    c = lambda t: t.detach().cpu().numpy()
    ret, cache = O.render_rnb(np_state(sdf), np_state(col), float(var.variance), c(b["rays_o"]), c(b["rays_d"]),
                              c(b["near"]), c(b["far"]), c(b["lights_dir"]), None, 1.0, warm, no_albedo, z_vals=c(z_vals))
    ref_loss, _, grads, _ = O.train_step_grads(ret, cache, c(b["true_rgb"]), c(b["mask"]), 0.1, 0.1)
    assert rel_l2(c(out["color_fine"]), ret["color_fine"]) < 1e-3
    assert rel_l2(c(out["weight_sum"]), ret["weight_sum"]) < 1e-3
    assert abs(float(loss) / ref_loss - 1) < 1e-3
    worst = 0.0
    for tag, mod in (("sdf", sdf), ("color", col)):
        for pname, p in sorted(mod.named_parameters()):
            key = f"{tag}.{pname}"
            if key not in grads:
                continue
            got, ref = c(p.grad).ravel(), np.asarray(grads[key]).ravel()
            if np.linalg.norm(ref) < 1e-12:
                continue
            single = float(np.abs(ref).max()) ** 2 > 0.99 * float((ref ** 2).sum())
            assert cosine(got, ref) > 0.999, (key, cosine(got, ref))
            check(f"oracle192[warm={warm},no_albedo={no_albedo}] gradient {key}{' (one-element tensor)' if single else ''}", rel_l2(got, ref), 5e-3)        # north_star 1e-2; measured worst 2.1e-3
            worst = max(worst, rel_l2(got, ref))
    assert abs(float(var.variance.grad) / float(grads["variance"]) - 1) < 1e-2
    print("worst per-tensor rel-L2:", worst)


def test_extract_fields_golden_and_oracle():
    from models.renderer import extract_fields
    g = load_golden("grid")
    _, sdf, _, _ = build_nets(True)
    bmin, bmax = torch.tensor(g["bmin"]), torch.tensor(g["bmax"])
    u = extract_fields(bmin, bmax, 32, sdf_network=sdf)
    assert u.shape == (32, 32, 32) and u.dtype == np.float32 and u.flags["C_CONTIGUOUS"]
    assert rel_l2(u, g["u32"]) < 1e-3
    # slabs tile the lattice exactly
    from rnb_b200 import grid
    parts = [grid.sdf_slab(sdf, bmin, bmax, 32, *grid.slab_bounds(32, r, 3)).cpu().numpy() for r in range(3)]
    assert np.array_equal(np.concatenate(parts, 0), u)
    # full-resolution lattice arithmetic: nodes of the 512^3 grid through a 3-wide slab at the far end of x
    idx = g["idx512"]
    sl = grid.sdf_slab(sdf, bmin, bmax, 512, 509, 512).cpu().numpy()        # x = 509..511
    ref = g["u512"][-3:]                                                     # idx512 ends with 509, 510, 511
    assert rel_l2(sl[:, idx][:, :, idx], ref) < 1e-3
    Ws, bs = O.sdf_effective(np_state(sdf))
    ref_slab = O.extract_fields(Ws, bs, g["bmin"], g["bmax"], 32, x_range=(4, 6))
    assert rel_l2(u[4:6], ref_slab) < 1e-3
