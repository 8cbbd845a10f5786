"""GPU parity of the per-ray kernels (sampling, compositing + adjoint) against the oracle / golden fixtures."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2, check
from oracle import rnb_oracle as O

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def test_searchsorted_indices_bit_exact():
    """Indices from searchsorted must be bit-exact given the same CDF (north_star)."""
    from rnb_b200 import kernels as K
    g = load_golden("sample_pdf")
    samples, inds = K.sample_pdf_from_cdf(cu(g["bins"]), cu(g["cdf"]), 16)
    assert np.array_equal(inds.cpu().numpy(), g["inds"])
    assert np.abs(samples.cpu().numpy() - g["samples"]).max() < 1e-6
    # random CDFs with ties and flat stretches, against the oracle
    rng = np.random.default_rng(3)
    w = rng.random((64, 111)).astype(np.float32) ** 6
    w[:, 40:60] = 0
    bins = np.sort(rng.random((64, 112)).astype(np.float32) * 2 + 1, -1)
    s_ref, i_ref, cdf = O.sample_pdf_det(bins, w, 16)
    samples, inds = K.sample_pdf_from_cdf(cu(bins), cu(cdf), 16)
    assert np.array_equal(inds.cpu().numpy(), i_ref)
    assert np.abs(samples.cpu().numpy() - s_ref).max() < 1e-6


@pytest.mark.parametrize("case", ["warmup_albedo", "init_warmup_albedo", "post_albedo"])
def test_upsample_steps(case):
    from rnb_b200 import kernels as K
    g = load_golden("render_" + case)
    o, d = cu(g["rays_o"]), cu(g["rays_d"])
    z0 = K.coarse_z(cu(g["near"]), cu(g["far"]), cu(g["t_rand"]), 64)
    assert np.abs(z0.cpu().numpy() - g["up0_z_in"]).max() < 1e-6
    for i in range(4):
        z_in, s_in = cu(g[f"up{i}_z_in"]), cu(g[f"up{i}_sdf_in"])
        z_new, _, _, inds, cdf = K.upsample_step(o, d, z_in, s_in, float(g[f"up{i}_inv_s"]), 16, want_debug=True)
        dz = np.abs(z_new.cpu().numpy() - g[f"up{i}_new_z"])
        # same conditioning caveat as the oracle test: fp32 cancellation where the pdf is flat
        assert dz.max() < 1e-3 and dz.mean() < 1e-4, (i, dz.max(), dz.mean())
        # own CDF -> oracle search reproduces the kernel's indices exactly
        _, i_ref, _ = O.sample_pdf_det(g[f"up{i}_z_in"], np.zeros_like(g[f"up{i}_z_in"][:, 1:]), 16,
                                       cdf_override=cdf.cpu().numpy())
        assert np.array_equal(inds.cpu().numpy(), i_ref)
    # merge path: step i+1 fed with pending samples equals the reference's merged arrays
    rng = np.random.default_rng(0)
    z_in, s_in = g["up1_z_in"], g["up1_sdf_in"]
    z_prev, s_prev = g["up0_z_in"], g["up0_sdf_in"]
    new_z = g["up0_new_z"]
    # the sdf of the pending samples = values the reference gathered into up1_sdf_in
    zm, sm = O.cat_z_vals(z_prev, new_z, s_prev, np.zeros_like(new_z))
    assert np.abs(zm - z_in).max() == 0
    new_sdf = np.stack([s_in[b][np.searchsorted(z_in[b], new_z[b])] for b in range(z_in.shape[0])])
    _, z_m, s_m, _, _ = K.upsample_step(o, d, cu(z_prev), cu(s_prev), 128.0, 16, cu(new_z), cu(new_sdf))
    assert np.array_equal(z_m.cpu().numpy(), z_in)
    assert np.abs(s_m.cpu().numpy() - s_in).max() < 1e-6
    z_f, mid = K.final_merge(cu(g["up3_z_in"]), cu(g["up3_new_z"]), 2.0 / 64)
    assert np.array_equal(z_f.cpu().numpy(), g["z_vals"])


@pytest.mark.parametrize("case", ["warmup_albedo", "warmup_noalbedo", "post_albedo", "post_noalbedo", "womask_anneal",
                                  "init_warmup_albedo"])
def test_composite_fwd_bwd(case):
    from rnb_b200 import kernels as K
    g = load_golden("render_" + case)
    B = g["z_vals"].shape[0]
    no_albedo = bool(g["no_albedo"])
    warm = bool(g["warmup"])
    grad = g["out_gradients"]
    albedo = np.ones_like(g["albedo"]) if no_albedo else g["albedo"]
    variance = 0.3 if case.startswith("init") else 0.45
    inv_s = float(np.exp(10 * np.float32(variance)))
    r = float(g["r"])
    fw = O.composite_forward(g["rays_o"], g["rays_d"], g["z_vals"], g["sdf"], grad, albedo, g["lights_dir"], inv_s, r, warm)
    var_t = torch.tensor(variance, device="cuda")
    p = K.composite_params(cu(g["rays_o"]), cu(g["rays_d"]), cu(g["z_vals"]), cu(g["sdf"]).view(-1), cu(grad).view(-1, 3),
                           None if no_albedo else cu(albedo).view(-1, 3), cu(g["lights_dir"]), var_t, r, 1 if warm else 0,
                           2.0 / 64)
    out = K.composite_fwd(p)
    for k_mine, k_ref in (("color", "color_fine"), ("weights", "weights"), ("cdf", "cdf_fine"), ("weight_sum", "weight_sum"),
                          ("weight_max", "weight_max")):
        assert rel_l2(out[k_mine].cpu().numpy(), g["out_" + k_ref]) < 1e-3, k_mine      # vs reference
        assert rel_l2(out[k_mine].cpu().numpy(), fw[k_ref]) < 1e-3, k_mine              # vs oracle
    assert np.array_equal(out["inside"].cpu().numpy(), g["out_inside_sphere"])
    ep = out["eik_part"].double().sum(0).cpu().numpy()
    assert abs(ep[0] / (ep[1] + 1e-5) / float(g["out_gradient_error"]) - 1) < 1e-4
    # adjoint with the loss' own cotangents
    _, _, (d_color, d_ws, d_eik) = O.rnb_loss(fw["color_fine"], fw["weight_sum"], fw["gradient_error"], g["true_rgb"],
                                              g["mask_used"], 0.1, float(g["mask_weight"]))
    rs, rg, ra, rinv = O.composite_backward(fw, g["rays_d"], g["sdf"], grad, albedo, g["lights_dir"], inv_s, r, warm,
                                            d_color, d_ws, d_eik)
    bw = K.composite_bwd(p, cu(d_color), cu(d_ws), torch.tensor(float(d_eik), device="cuda"),
                         torch.tensor(float(fw["relax"].sum()), device="cuda"), not no_albedo)
    check(f"compositing adjoint[{case}] d_sdf vs oracle", rel_l2(bw["d_sdf"].cpu().numpy(), rs), 1e-3)
    check(f"compositing adjoint[{case}] d_grad vs oracle", rel_l2(bw["d_grad"].cpu().numpy(), rg.reshape(-1, 3)), 1e-3)
    if not no_albedo:
        check(f"compositing adjoint[{case}] d_albedo vs oracle", rel_l2(bw["d_albedo"].cpu().numpy(), ra.reshape(-1, 3)), 1e-3)
    dvar = float(bw["d_var_part"].double().sum())
    check(f"compositing adjoint[{case}] d_variance vs oracle", abs(dvar / (rinv * 10 * inv_s) - 1), 1e-3)
