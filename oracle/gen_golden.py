"""Generate tests/golden/*.npz by RUNNING THE REFERENCE (build container only).

TEST INFRASTRUCTURE.  Usage:  python oracle/gen_golden.py
The reference holds no tests or golden vectors, so the pins are its own
outputs on seeded synthetic inputs: weights from `torch.manual_seed(0)` (+ a seeded
perturbation for "trained-like" cases), inputs from rnb_b200.synth.make_batch.  The jitter
`torch.rand([B,1])` (models/renderer.py:844, 948) is injected so CPU and GPU runs see the
same value.  Parameter gradients are stored sub-sampled (every GRAD_STRIDE-th element)
together with their full L2 norm to keep the fixtures small.
"""
from __future__ import annotations

import os
import sys
from contextlib import contextmanager

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rnb-neus-fork_b200"))

from oracle import ref_loader  # noqa: E402
from rnb_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
GRAD_STRIDE = 5


def np32(t):
    return t.detach().cpu().numpy().astype(np.float32)


def build_reference_nets(perturb=False, with_nerf=False):
    ref = ref_loader.load()
    torch.manual_seed(0)
    conf = synth.WMASK_CONF
    nerf = ref.fields.NeRF(**conf["nerf"])
    sdf = ref.fields.SDFNetwork(**conf["sdf_network"])
    var = ref.fields.SingleVarianceNetwork(**conf["variance_network"])
    col = ref.fields.RenderingNetwork(**conf["rendering_network"])
    if perturb:
        synth.perturb_state_dict_(sdf, synth.SDF_NOISE, 5)
        synth.perturb_state_dict_(col, synth.COLOR_NOISE, 6)
        with torch.no_grad():
            var.variance.fill_(synth.TRAINED_VARIANCE)
    return ref, nerf, sdf, var, col


@contextmanager
def injected_rand(values):
    """Make torch.rand(shape) return the queued tensors (consumed in order)."""
    queue = list(values)
    orig = torch.rand

    def fake(*a, **k):
        t = queue.pop(0)
        shape = list(a[0]) if len(a) == 1 and isinstance(a[0], (list, tuple)) else list(a)
        assert list(t.shape) == shape, (t.shape, shape)
        return t.clone()

    torch.rand = fake
    try:
        yield
    finally:
        torch.rand = orig


def sub(t):
    return np32(t).reshape(-1)[::GRAD_STRIDE].copy()


def weight_checksum(module):
    return np.array([float(p.detach().double().sum()) for _, p in sorted(module.named_parameters())])


def gen_embed():
    ref = ref_loader.load()
    g = torch.Generator().manual_seed(11)
    x = (torch.rand(16, 3, generator=g) - 0.5) * 2.4
    e6, d6 = ref.embedder.get_embedder(6, 3)
    e4, d4 = ref.embedder.get_embedder(4, 3)
    x4 = (torch.rand(8, 4, generator=g) - 0.5) * 2.0
    e10, d10 = ref.embedder.get_embedder(10, 4)
    np.savez(os.path.join(GOLD, "embed.npz"), x=np32(x), e6=np32(e6(x)), e4=np32(e4(x)), x4=np32(x4),
             e10=np32(e10(x4)), dims=np.array([d6, d4, d10]))


def gen_sdf(perturb):
    ref, nerf, sdf, var, col = build_reference_nets(perturb)
    g = torch.Generator().manual_seed(12)
    x = (torch.rand(96, 3, generator=g) - 0.5) * 2.2
    x[:8] *= 0.05                                  # points near the origin
    x[8:16] = x[8:16] / x[8:16].norm(dim=-1, keepdim=True) * 0.5   # near the zero level set of the init sphere
    xr = x.clone().requires_grad_(True)
    out = sdf(xr)
    grad = sdf.gradient(xr).squeeze(1)
    ybar = torch.randn(out.shape, generator=g) * torch.tensor([1.0] + [1e-3] * 256)
    gbar = torch.randn(grad.shape, generator=g) * 0.1
    sdf.zero_grad()
    ((out * ybar).sum() + (grad * gbar).sum()).backward()
    d = dict(x=np32(x), out=np32(out), grad=np32(grad), ybar=np32(ybar), gbar=np32(gbar),
             wsum=weight_checksum(sdf), stride=np.array(GRAD_STRIDE))
    for name, p in sorted(sdf.named_parameters()):
        d["g_" + name] = sub(p.grad)
        d["n_" + name] = np.array(float(p.grad.double().norm()))
    # albedo net on the same points
    feat = out[:, 1:].detach()
    nrm = grad.detach()
    nr = nrm.clone().requires_grad_(True)
    fr = feat.clone().requires_grad_(True)
    alb = col(x, nr, nr, fr)
    abar = torch.randn(alb.shape, generator=g)
    col.zero_grad()
    (alb * abar).sum().backward()
    d.update(albedo=np32(alb), abar=np32(abar), d_normals=np32(nr.grad), d_feat=np32(fr.grad),
             cwsum=weight_checksum(col))
    for name, p in sorted(col.named_parameters()):
        d["cg_" + name] = sub(p.grad)
        d["cn_" + name] = np.array(float(p.grad.double().norm()))
    np.savez_compressed(os.path.join(GOLD, f"sdf_{'perturbed' if perturb else 'init'}.npz"), **d)


def loss_fn(out, true_rgb, mask, igr_weight, mask_weight, n_lights):
    # exp_runner.py:241-256 restated (the Runner itself needs pyhocon/trimesh/CUDA)
    mask_sum = mask.sum() + 1e-5
    color_error = ((out["color_fine"] - true_rgb) * mask[None, :, :]).reshape(-1, 3)
    color_loss = F.l1_loss(color_error, torch.zeros_like(color_error), reduction="sum") / (mask_sum * n_lights)
    mask_loss = F.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), mask)
    return color_loss + out["gradient_error"] * igr_weight + mask_loss * mask_weight, color_loss, mask_loss


def gen_render(name, warmup, no_albedo, B=16, perturb=True, r=1.0, mask_weight=0.1, seed=1):
    ref, nerf, sdf, var, col = build_reference_nets(perturb)
    conf = synth.WMASK_CONF["neus_renderer"]
    renderer = ref.renderer.NeuSRenderer(nerf, sdf, var, col, **conf)
    renderer.color_depth = 3
    b = synth.make_batch(B, 3, warmup, seed)
    mask = b["mask"] if mask_weight > 0 else torch.ones_like(b["mask"])
    fn = renderer.render_rnb_warmup if warmup else renderer.render_rnb
    # capture per-step up-sampling tensors
    steps = []
    orig_up, orig_cat = renderer.up_sample, renderer.cat_z_vals

    def up(rays_o, rays_d, z_vals, sdf_v, n_imp, inv_s):
        new_z = orig_up(rays_o, rays_d, z_vals, sdf_v, n_imp, inv_s)
        steps.append(dict(z_in=np32(z_vals), sdf_in=np32(sdf_v.reshape(z_vals.shape)), new_z=np32(new_z),
                          inv_s=inv_s))
        return new_z

    renderer.up_sample = up
    captured = {}
    orig_core = renderer.render_core_mvps

    def core(rays_o, rays_d, z_vals, *a, **k):
        captured["z_vals"] = np32(z_vals)
        ret = orig_core(rays_o, rays_d, z_vals, *a, **k)
        captured["sdf"] = np32(ret["sdf"]).reshape(z_vals.shape)
        captured["albedo"] = np32(ret["sampled_albedo"])
        return ret

    renderer.render_core_mvps = core
    for m in (sdf, var, col):
        m.zero_grad()
    with injected_rand([b["t_rand"] + 0.5]):
        out = fn(b["rays_o"], b["rays_d"], b["near"], b["far"], b["lights_dir"], cos_anneal_ratio=r,
                 no_albedo=no_albedo)
    loss, closs, mloss = loss_fn(out, b["true_rgb"], mask, 0.1, mask_weight, 3)
    loss.backward()
    d = {k: np32(v) for k, v in b.items()}
    d["mask_used"] = np32(mask)
    d.update(z_vals=captured["z_vals"], sdf=captured["sdf"], albedo=captured["albedo"],
             loss=np.array(float(loss)), color_loss=np.array(float(closs)), mask_loss=np.array(float(mloss)),
             r=np.array(r), mask_weight=np.array(mask_weight), warmup=np.array(warmup), no_albedo=np.array(no_albedo),
             stride=np.array(GRAD_STRIDE))
    for k in ("color_fine", "s_val", "cdf_fine", "weight_sum", "weight_max", "gradients", "weights",
              "gradient_error", "inside_sphere"):
        d["out_" + k] = np32(out[k])
    for i, s in enumerate(steps):
        for k, v in s.items():
            d[f"up{i}_{k}"] = np.asarray(v)
    for tag, mod in (("sdf", sdf), ("color", col), ("var", var)):
        for pname, p in sorted(mod.named_parameters()):
            if p.grad is None:
                continue
            d[f"g_{tag}.{pname}"] = sub(p.grad)
            d[f"n_{tag}.{pname}"] = np.array(float(p.grad.double().norm()))
    np.savez_compressed(os.path.join(GOLD, f"render_{name}.npz"), **d)
    print(name, "loss", float(loss), "eik", float(out["gradient_error"]))


LARGE_STRIDE = 16


def gen_render_large(name, warmup, no_albedo, B=512, r=1.0, mask_weight=0.1, seed=21):
    """BASELINE.json configs 0-2 size (512 rays, 65 536 fine points): the reference's outputs, loss and parameter
    gradients (every LARGE_STRIDE-th element + the full per-tensor norms) AND the float64 oracle's gradients on the same
    inputs and the reference's own sample depths -- so the GPU test at this size is pinned to the reference, and the
    oracle is pinned to the reference at this size as well (tests/test_oracle_golden.py)."""
    from oracle import rnb_oracle as O
    ref, nerf, sdf, var, col = build_reference_nets(True)
    renderer = ref.renderer.NeuSRenderer(nerf, sdf, var, col, **synth.WMASK_CONF["neus_renderer"])
    renderer.color_depth = 3
    b = synth.make_batch(B, 3, warmup, seed)
    mask = b["mask"] if mask_weight > 0 else torch.ones_like(b["mask"])
    fn = renderer.render_rnb_warmup if warmup else renderer.render_rnb
    captured = {}
    orig_core = renderer.render_core_mvps

    def core(rays_o, rays_d, z_vals, *a, **k):
        captured["z_vals"] = np32(z_vals)
        return orig_core(rays_o, rays_d, z_vals, *a, **k)

    renderer.render_core_mvps = core
    for m in (sdf, var, col):
        m.zero_grad()
    with injected_rand([b["t_rand"] + 0.5]):
        out = fn(b["rays_o"], b["rays_d"], b["near"], b["far"], b["lights_dir"], cos_anneal_ratio=r, no_albedo=no_albedo)
    loss, closs, mloss = loss_fn(out, b["true_rgb"], mask, 0.1, mask_weight, 3)
    loss.backward()
    d = dict(B=np.array(B), seed=np.array(seed), t_rand=np32(b["t_rand"]), mask_used=np32(mask), z_vals=captured["z_vals"],
             loss=np.array(float(loss)), r=np.array(r), mask_weight=np.array(mask_weight), warmup=np.array(warmup),
             no_albedo=np.array(no_albedo), stride=np.array(LARGE_STRIDE))
    for k in ("color_fine", "weight_sum", "weight_max", "s_val", "gradient_error"):
        d["out_" + k] = np32(out[k])
    d["out_gradients_head"] = np32(out["gradients"][:32])        # normals of the first 32 rays
    d["out_weights_head"] = np32(out["weights"][:32])
    lsub = lambda t: np.asarray(t, np.float64).reshape(-1)[::LARGE_STRIDE].astype(np.float32)
    for tag, mod in (("sdf", sdf), ("color", col), ("var", var)):
        for pname, p in sorted(mod.named_parameters()):
            if p.grad is None:
                continue
            d[f"g_{tag}.{pname}"] = lsub(p.grad.detach().numpy())
            d[f"n_{tag}.{pname}"] = np.array(float(p.grad.double().norm()))
    # the float64 oracle on the same inputs, the same jitter-free sample depths
    sd64 = lambda m: {k: v.detach().double().numpy() for k, v in m.state_dict().items()}
    c = lambda t: t.detach().numpy()
    ret, cache = O.render_rnb(sd64(sdf), sd64(col), float(var.variance), c(b["rays_o"]), c(b["rays_d"]), c(b["near"]),
                              c(b["far"]), c(b["lights_dir"]), None, r, warmup, no_albedo, z_vals=captured["z_vals"].astype(np.float64))
    o_loss, _, grads, _ = O.train_step_grads(ret, cache, c(b["true_rgb"]), c(mask), 0.1, mask_weight)
    d["o_loss"] = np.array(float(o_loss))
    d["o_color_fine"] = np.asarray(ret["color_fine"], np.float32)
    d["o_weight_sum"] = np.asarray(ret["weight_sum"], np.float32)
    d["o_gradient_error"] = np.array(float(ret["gradient_error"]))
    for key, gval in grads.items():
        d["og_" + key] = lsub(gval)
        d["on_" + key] = np.array(float(np.linalg.norm(np.asarray(gval, np.float64))))
    np.savez_compressed(os.path.join(GOLD, f"render512_{name}.npz"), **d)
    print("large", name, "loss", float(loss), "oracle loss", float(o_loss))


def gen_large():
    gen_render_large("warmup_albedo", True, False)
    gen_render_large("warmup_noalbedo", True, True)
    gen_render_large("post_albedo", False, False)
    gen_render_large("post_noalbedo", False, True)
    gen_render_large("womask_anneal", False, False, r=0.3, mask_weight=0.0, seed=22)


def gen_upsample_search():
    """searchsorted pin: cdf -> inds, samples (models/renderer.py:39-69)."""
    ref = ref_loader.load()
    g = torch.Generator().manual_seed(13)
    B, n = 12, 80
    bins = torch.sort(torch.rand(B, n, generator=g) * 2 + 2, -1)[0]
    w = torch.rand(B, n - 1, generator=g) ** 4
    w[0] = 0.0                 # all-zero weights -> uniform pdf
    w[1, :] = 0.0
    w[1, 17] = 1.0             # a single spike: denom < 1e-5 branches elsewhere
    captured = {}
    orig = torch.searchsorted

    def ss(cdf, u, right=False):
        captured["cdf"] = np32(cdf)
        r = orig(cdf, u, right=right)
        captured["inds"] = r.numpy().copy()
        return r

    torch.searchsorted = ss
    try:
        samples = ref.renderer.sample_pdf(bins, w, 16, det=True)
    finally:
        torch.searchsorted = orig
    np.savez(os.path.join(GOLD, "sample_pdf.npz"), bins=np32(bins), weights=np32(w), samples=np32(samples),
             cdf=captured["cdf"], inds=captured["inds"])


def gen_grid():
    ref, nerf, sdf, var, col = build_reference_nets(True)
    bmin = torch.tensor([-1.01, -1.01, -1.01])
    bmax = torch.tensor([1.01, 1.01, 1.01])
    with torch.no_grad():
        u32 = ref.renderer.extract_fields(bmin, bmax, 32, lambda p: -sdf.sdf(p))
        # corner 12^3 block of the 512^3 grid pins the linspace arithmetic at full resolution
        X = torch.linspace(bmin[0], bmax[0], 512)
        idx = torch.tensor([0, 1, 2, 100, 255, 256, 257, 300, 509, 510, 511])
        xs = X[idx]
        xx, yy, zz = torch.meshgrid(xs, xs, xs, indexing="ij")
        pts = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], -1)
        u512 = (-sdf.sdf(pts)).reshape(len(idx), len(idx), len(idx))
    np.savez_compressed(os.path.join(GOLD, "grid.npz"), bmin=np32(bmin), bmax=np32(bmax), u32=u32.astype(np.float32),
                        idx512=idx.numpy(), u512=np32(u512), axis512=np32(X), wsum=weight_checksum(sdf))


def gen_background(B=8):
    """NeRF++ background through render() with n_outside=32 (models/renderer.py:556-648) --
    render_rnb* raise when n_outside > 0, so this is the only runnable pin."""
    ref, nerf, sdf, var, col = build_reference_nets(True)
    synth.perturb_state_dict_(nerf, 0.02, 7)
    conf = dict(synth.WMASK_CONF["neus_renderer"], n_outside=32)
    # plain render() feeds view dirs; the shipped colour net is 'no_view_dir' with d_in=6 and works as is
    renderer = ref.renderer.NeuSRenderer(nerf, sdf, var, col, **conf)
    renderer.color_depth = 3
    b = synth.make_batch(B, 3, True, 3)
    g = torch.Generator().manual_seed(14)
    r2 = torch.rand(B, 32, generator=g)
    cap = {}
    core, outside = renderer.render_core, renderer.render_core_outside

    def core_hook(rays_o, rays_d, z_vals, *a, **k):
        cap["z_vals"] = z_vals.detach().clone()
        return core(rays_o, rays_d, z_vals, *a, **k)

    def outside_hook(rays_o, rays_d, z_vals, *a, **k):
        cap["z_feed"] = z_vals.detach().clone()
        ret = outside(rays_o, rays_d, z_vals, *a, **k)
        cap["bg_alpha"], cap["bg_color"] = ret["alpha"].detach().clone(), ret["sampled_color"].detach().clone()
        return ret

    renderer.render_core, renderer.render_core_outside = core_hook, outside_hook
    with injected_rand([b["t_rand"] + 0.5, r2]):
        out = renderer.render(b["rays_o"], b["rays_d"], b["near"], b["far"], cos_anneal_ratio=1.0,
                              background_rgb=None)
    d = {k: np32(v) for k, v in b.items()}
    d["rand_outside"] = np32(r2)
    for k, v in cap.items():
        d[k] = np32(v)
    for k in ("weight_max", "cdf_fine", "gradients", "s_val"):
        d["out_" + k] = np32(out[k])
    for k in ("color_fine", "weights", "weight_sum", "inside_sphere", "gradient_error"):
        d["out_" + k] = np32(out[k])
    # direct NeRF pin
    pts4 = torch.randn(32, 4, generator=g)
    pts4 = pts4 / pts4[:, :3].norm(dim=-1, keepdim=True).clamp(min=1.0)
    dirs = torch.randn(32, 3, generator=g)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    with torch.no_grad():
        a, rgb = nerf(pts4, dirs)
    d.update(nerf_pts=np32(pts4), nerf_dirs=np32(dirs), nerf_alpha=np32(a), nerf_rgb=np32(rgb),
             nwsum=weight_checksum(nerf))
    np.savez_compressed(os.path.join(GOLD, "background.npz"), **d)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    if "--large" in sys.argv:          # only the 512-ray fixtures (about a minute of oracle time each)
        gen_large()
        return
    gen_embed()
    gen_sdf(False)
    gen_sdf(True)
    gen_upsample_search()
    gen_render("warmup_albedo", True, False)
    gen_render("warmup_noalbedo", True, True)
    gen_render("post_albedo", False, False)
    gen_render("post_noalbedo", False, True)
    gen_render("womask_anneal", False, False, r=0.3, mask_weight=0.0, seed=2)
    gen_render("init_warmup_albedo", True, False, perturb=False, seed=4)
    gen_grid()
    gen_background()
    gen_large()
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
