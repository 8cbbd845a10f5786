"""Short single-GPU program for ncu captures: two train_rnb steps (albedo on) at --rays rays."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rnb-neus-fork_b200")]
import torch
from bench import build, loss_fn
from rnb_b200 import synth
ap = argparse.ArgumentParser(); ap.add_argument("--rays", type=int, default=1024); ap.add_argument("--steps", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda", 0)
renderer, sdf, var, col = build(dev)
b = {k: v.to(dev) for k, v in synth.make_batch(a.rays, 3, True, 1).items()}
for i in range(a.steps):
    for m in (sdf, var, col):
        m.zero_grad()
    out = renderer.render_rnb_warmup(b["rays_o"], b["rays_d"], b["near"], b["far"], b["lights_dir"], cos_anneal_ratio=1.0)
    loss_fn(out, b["true_rgb"], b["mask"]).backward()
torch.cuda.synchronize()
print("done", float(out["gradient_error"]))
