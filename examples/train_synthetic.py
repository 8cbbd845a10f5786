"""End-to-end run on an analytic scene: train_rnb loop -> mesh -> PLY, no data files needed (needs a B200).

    python examples/train_synthetic.py --iters 2000 --batch 512 --out /tmp/sphere.ply

Mirrors what `python exp_runner.py --mode train_rnb ...` followed by `--mode validate_mesh` does in the reference
(exp_runner.py:147-303, 560-580) on a sphere of radius 0.6 seen by 8 cameras under 3 lights.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "rnb-neus-fork_b200")]

import numpy as np  # noqa: E402
import torch  # noqa: E402

from models.fields import NeRF, RenderingNetwork, SDFNetwork, SingleVarianceNetwork  # noqa: E402
from models.renderer import NeuSRenderer  # noqa: E402
from rnb_b200 import synth  # noqa: E402
from rnb_b200.meshio import write_ply  # noqa: E402
from rnb_b200.scene import sphere_scene  # noqa: E402
from rnb_b200.train_loop import train_rnb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--warmup-iters", type=int, default=200, help="iterations of render_rnb_warmup (conf train.warm_up_iter)")
    ap.add_argument("--radius", type=float, default=0.6)
    ap.add_argument("--res", type=int, default=128, help="marching-cubes resolution")
    ap.add_argument("--out", default="")
    ap.add_argument("--graph", action="store_true", help="one CUDA graph per step (launch-bound 512-ray regime)")
    a = ap.parse_args()
    dev = torch.device("cuda")
    torch.manual_seed(0)
    conf = synth.WMASK_CONF
    nerf = NeRF(**conf["nerf"]).to(dev)
    sdf = SDFNetwork(**conf["sdf_network"]).to(dev)
    var = SingleVarianceNetwork(**conf["variance_network"]).to(dev)
    col = RenderingNetwork(**conf["rendering_network"]).to(dev)
    renderer = NeuSRenderer(nerf, sdf, var, col, **conf["neus_renderer"])
    scene = sphere_scene(radius=a.radius)
    marks = {}

    def log(msg, it=None):
        print(msg)

    def on_iter(it):                       # steady-state clock: skip the first 10 % (lazy initialisation, allocator growth)
        if it == a.iters // 10:
            torch.cuda.synchronize()
            marks["t0"], marks["it0"] = time.time(), it
    _, hist = train_rnb(renderer, [sdf, var, col], scene, a.iters, batch_size=a.batch, warm_up_iter=a.warmup_iters,
                        report_freq=max(1, a.iters // 10), log=log, on_iter=on_iter, use_graph=a.graph)
    torch.cuda.synchronize()
    dt, n = time.time() - marks["t0"], a.iters - marks["it0"]
    print(f"{n} iterations x {a.batch} rays in {dt:.2f} s = {1e3 * dt / n:.2f} ms/iteration, {n * a.batch / dt:.3e} rays/s "
          "(whole loop: batch gather, render, loss, backward, Adam, host code)")
    v, t = renderer.extract_geometry(torch.tensor([-1.01] * 3), torch.tensor([1.01] * 3), resolution=a.res, threshold=0.0)
    r = np.linalg.norm(v, axis=1)
    print(f"mesh: {len(v)} vertices, {len(t)} triangles, radius {r.mean():.4f} +- {r.std():.4f} (target {a.radius})")
    if a.out:
        write_ply(a.out, v, t)
        print("wrote", a.out)


if __name__ == "__main__":
    main()
