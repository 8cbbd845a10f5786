import sys; sys.path[:0]=['.','tests','rnb-neus-fork_b200']
import numpy as np, torch
from conftest import load_golden, rel_l2, cosine
from test_gpu_e2e import *
for case in sys.argv[1:]:
    g = load_golden("render_" + case)
    renderer, sdf, var, col = make_renderer(not case.startswith("init"))
    warm, no_albedo = bool(g["warmup"]), bool(g["no_albedo"])
    args = (cu(g["rays_o"]), cu(g["rays_d"]), cu(g["near"]), cu(g["far"]), cu(g["lights_dir"]))
    out = renderer._render_rnb(warm, *args, -1, None, float(g["r"]), no_albedo, _z_vals=cu(g["z_vals"]))
    loss = loss_fn(out, cu(g["true_rgb"]), cu(g["mask_used"]), float(g["mask_weight"]))
    loss.backward()
    st = int(g["stride"])
    print(case, "loss", float(loss), float(g["loss"]))
    for tag, mod in (("sdf", sdf), ("color", col), ("var", var)):
        for pname, p in sorted(mod.named_parameters()):
            key = f"g_{tag}.{pname}"
            if key not in g or p.grad is None: continue
            got = p.grad.detach().cpu().numpy().reshape(-1); ref = g[key]
            print(f"  {key:28s} cos {cosine(got[::st], ref):.6f} rel {rel_l2(got[::st], ref):.2e} norm {np.linalg.norm(ref):.3e}")
