"""Diagnostic: per-parameter gradient difference of one train step between torch's per-layer weight-norm fold and the fused
fold (RNB_FUSED_WN=0 / 1), next to the sensitivity floor: the per-layer fold with every weight_g moved by one fp32 ulp.
Run on a GPU box: python profiles/_wn_diag.py [n_rays]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "rnb-neus-fork_b200"))
import torch

from rnb_b200 import synth
from test_gpu_e2e import loss_fn, make_renderer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
renderer, sdf, var, col = make_renderer(True)
renderer.perturb = 0.0
named = [(mn + "." + n, p) for mn, m in (("sdf", sdf), ("var", var), ("col", col)) for n, p in m.named_parameters()]
params = [p for _, p in named]
b = {k: v.cuda() for k, v in synth.make_batch(B, 3, True, 3).items()}


def step(flag):
    os.environ["RNB_FUSED_WN"] = flag
    for p in params:
        p.grad = None
    torch.manual_seed(11)
    out = renderer.render_rnb_warmup(b["rays_o"], b["rays_d"], b["near"], b["far"], b["lights_dir"], cos_anneal_ratio=1.0)
    loss = loss_fn(out, b["true_rgb"], b["mask"], 0.1)
    loss.backward()
    return float(loss.detach()), [p.grad.clone() for p in params]


l0, g0 = step("0")
l1, g1 = step("1")
with torch.no_grad():
    for n, p in named:
        if n.endswith("weight_g"):
            p.copy_(torch.nextafter(p, p * 2))
l2, g2 = step("0")
print("loss torch %.9g fused %.9g torch+1ulp %.9g" % (l0, l1, l2))
print("%-22s %10s %12s %12s" % ("param", "|g|", "fused/torch", "1ulp/torch"))
for (n, _), a, f, u in zip(named, g0, g1, g2):
    na = float(a.norm())
    print("%-22s %10.3e %12.3e %12.3e" % (n, na, float((a - f).norm()) / (na + 1e-30), float((a - u).norm()) / (na + 1e-30)))
