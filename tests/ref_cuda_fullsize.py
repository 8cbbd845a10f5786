"""Runs the UNMODIFIED reference (models/renderer.py:828-930 render_rnb_warmup / 932-1033 render_rnb + the loss of exp_runner.py:240-260) through
PyTorch-CUDA, fp32 with TF32 off, on the benchmark's own batch -- 8192 rays, 1 048 576 fine points, "trained-like" perturbed
weights -- and writes its outputs, its own sample depths, the loss and every parameter gradient to an npz.  Started as a
subprocess by tests/test_gpu_fullsize.py: the reference needs the process-wide
torch.set_default_tensor_type('torch.cuda.FloatTensor') of exp_runner.py:669."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rnb-neus-fork_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", required=True)
ap.add_argument("--rays", type=int, default=8192)
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--warm", type=int, default=1)
ap.add_argument("--no-albedo", type=int, default=0)
ap.add_argument("--bg", type=int, default=0, help="instead of a train step: render() with the NeRF++ background, n_outside = 32")
ap.add_argument("--grid", type=int, default=0, help="instead of a train step: extract_fields at this resolution -> .npy")
args = ap.parse_args()

from oracle.gen_golden import build_reference_nets, injected_rand, loss_fn  # noqa: E402
from rnb_b200 import synth  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
ref, nerf, sdf, var, col = build_reference_nets(True)
for m in (nerf, sdf, var, col):
    m.to(dev)
renderer = ref.renderer.NeuSRenderer(nerf, sdf, var, col, **synth.WMASK_CONF["neus_renderer"])
renderer.color_depth = 3
if args.grid:
    # validate_mesh's lattice (exp_runner.py:561-578 -> models/renderer.py:10-25, 1219-1224) with the reference's own code
    import time
    torch.set_default_tensor_type("torch.cuda.FloatTensor")
    bmin = torch.tensor([-1.01, -1.01, -1.01])
    bmax = torch.tensor([1.01, 1.01, 1.01])
    torch.cuda.synchronize()
    t0 = time.time()
    with torch.device(dev):
        u = ref.renderer.extract_fields(bmin, bmax, args.grid, lambda pts: -sdf.sdf(pts))
    dt = time.time() - t0
    np.save(args.out, u)
    print("REF_DONE grid", args.grid, "seconds", round(dt, 2))
    sys.exit(0)
if args.bg:
    # NeuSRenderer.render() with n_outside = 32 (models/renderer.py:556-648), forward only like its one reference caller
    synth.perturb_state_dict_(nerf, 0.02, 7)
    conf = dict(synth.WMASK_CONF["neus_renderer"], n_outside=32)
    renderer = ref.renderer.NeuSRenderer(nerf, sdf, var, col, **conf)
    renderer.color_depth = 3
    b = {k: v.to(dev) for k, v in synth.make_batch(args.rays, 3, True, args.seed).items()}
    r2 = torch.rand(args.rays, 32, generator=torch.Generator().manual_seed(14)).to(dev)
    cap = {}
    core0, outside0 = renderer.render_core, renderer.render_core_outside

    def core_hook(rays_o, rays_d, z_vals, *a, **k):
        cap["z_vals"] = z_vals.detach().clone()
        return core0(rays_o, rays_d, z_vals, *a, **k)

    def outside_hook(rays_o, rays_d, z_vals, *a, **k):
        cap["z_feed"] = z_vals.detach().clone()
        return outside0(rays_o, rays_d, z_vals, *a, **k)

    renderer.render_core, renderer.render_core_outside = core_hook, outside_hook
    torch.set_default_tensor_type("torch.cuda.FloatTensor")
    with torch.device(dev), injected_rand([b["t_rand"] + 0.5, r2]):       # (render() differentiates the SDF itself: no no_grad)
        out = renderer.render(b["rays_o"], b["rays_d"], b["near"], b["far"], cos_anneal_ratio=1.0, background_rgb=None)
    torch.cuda.synchronize()
    c = lambda t: t.detach().float().cpu().numpy()
    d = dict(seed=np.array(args.seed), rand_outside=c(r2), z_vals=c(cap["z_vals"]), z_feed=c(cap["z_feed"]))
    for k in ("color_fine", "weight_sum", "weight_max", "weights", "cdf_fine", "s_val", "gradient_error", "inside_sphere"):
        d["out_" + k] = c(out[k])
    d["out_gradients_head"] = c(out["gradients"][:64])
    np.savez(args.out, **d)
    print("REF_DONE render() with background,", args.rays, "rays")
    sys.exit(0)
warm, no_albedo = bool(args.warm), bool(args.no_albedo)
b = {k: v.to(dev) for k, v in synth.make_batch(args.rays, 3, warm, args.seed).items()}
captured = {}
orig_core = renderer.render_core_mvps


def core(rays_o, rays_d, z_vals, *a, **k):
    captured["z_vals"] = z_vals.detach().clone()
    return orig_core(rays_o, rays_d, z_vals, *a, **k)


renderer.render_core_mvps = core
torch.set_default_tensor_type("torch.cuda.FloatTensor")
with torch.device(dev), injected_rand([b["t_rand"] + 0.5]):
    fn = renderer.render_rnb_warmup if warm else renderer.render_rnb
    out = fn(b["rays_o"], b["rays_d"], b["near"], b["far"], b["lights_dir"], cos_anneal_ratio=1.0, no_albedo=no_albedo)
    loss = loss_fn(out, b["true_rgb"], b["mask"], 0.1, 0.1, 3)[0]
    loss.backward()
torch.cuda.synchronize()
c = lambda t: t.detach().float().cpu().numpy()
d = dict(rays=np.array(args.rays), seed=np.array(args.seed), loss=np.array(float(loss)), z_vals=c(captured["z_vals"]),
         peak_mem_gb=np.array(torch.cuda.max_memory_allocated(dev) / 2 ** 30))
for k in ("color_fine", "weight_sum", "weight_max", "s_val", "gradient_error"):
    d["out_" + k] = c(out[k])
d["out_gradients_head"] = c(out["gradients"][:64])
for tag, mod in (("sdf", sdf), ("color", col), ("var", var)):
    for pname, p in sorted(mod.named_parameters()):
        if p.grad is not None:
            d[f"g_{tag}.{pname}"] = c(p.grad)
np.savez(args.out, **d)
print("REF_DONE loss", float(loss), "peak GB", float(d["peak_mem_gb"]))
