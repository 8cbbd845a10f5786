"""torchrun --nproc-per-node N tests/_mesh_dist_check.py : sharded validate_mesh (lattice slab + marching cubes per GPU,
triangles gathered) against the single-GPU extraction; prints timings.  Manual multi-GPU check (NCCL)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rnb-neus-fork_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import torch
import torch.distributed as dist
from gpu_common import build_nets
from rnb_b200 import grid

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
_, sdf, _, _ = build_nets(True, device=f"cuda:{local}")
bmin, bmax = torch.tensor([-1.01] * 3), torch.tensor([1.01] * 3)
for R in (128, 512):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.time()
    out = grid.extract_mesh_distributed(sdf, bmin, bmax, R, 0.0)
    torch.cuda.synchronize(); dist.barrier(); t1 = time.time()
    if rank == 0:
        V, T = out
        u = grid.sdf_slab(sdf, bmin, bmax, R, 0, R)
        Vs, Ts = grid.marching_cubes_device(u, 0.0)
        Vs = Vs / (R - 1.0) * 2.02 - 1.01
        ok = V.shape == Vs.shape and T.shape == Ts.shape and np.allclose(np.sort(V, axis=0), np.sort(Vs, axis=0), atol=1e-6)
        print(f"R={R} world={world}: {len(T)} triangles, {len(V)} vertices, sharded == single-GPU: {ok}, {1e3 * (t1 - t0):.1f} ms end to end", flush=True)
        assert ok
dist.destroy_process_group()
