"""SDF lattice query for validate_mesh (reference models/renderer.py:10-36, 1219-1224).

`u[x,y,z] = -sdf(X[x], Y[y], Z[z])`, float32, C-contiguous.  The grid kernel generates the lattice points from
(bound_min, bound_max, R, slab) in registers and writes -sdf straight into a device slab: no meshgrid, no 64^3 block
loop, one D2H copy.  Slabs are contiguous in x (u[x0:x1]), so an N-GPU run is N independent launches and one gather.
"""
from __future__ import annotations

import numpy as np
import torch

from . import kernels as K
from . import lib as L
from . import ops


@torch.no_grad()
def sdf_slab(sdf_network, bound_min, bound_max, resolution, x0, x1, out=None):
    """Device tensor [x1-x0, R, R] = -sdf on the x-slab [x0, x1) of the R^3 lattice."""
    pk = ops.packed_sdf_nograd(sdf_network)
    bmin = [float(v) for v in torch.as_tensor(bound_min).detach().cpu().reshape(-1)]
    bmax = [float(v) for v in torch.as_tensor(bound_max).detach().cpu().reshape(-1)]
    nx = x1 - x0
    if out is None:
        out = torch.empty(nx, resolution, resolution, dtype=torch.float32, device=pk.wblob.device)
    K.sdf_fwd(pk, K.points_grid(bmin, bmax, resolution, x0, nx), out=out.view(-1), out_scale=-1.0)
    return out


def slab_bounds(resolution, rank, world):
    per = (resolution + world - 1) // world
    return min(resolution, rank * per), min(resolution, (rank + 1) * per)


@torch.no_grad()
def extract_fields(sdf_network, bound_min, bound_max, resolution):
    """Single-process path: the whole lattice on the current device -> numpy [R,R,R]."""
    return sdf_slab(sdf_network, bound_min, bound_max, resolution, 0, resolution).cpu().numpy()


@torch.no_grad()
def extract_fields_distributed(sdf_network, bound_min, bound_max, resolution, group=None):
    """z-slab sharding of the north_star, stored x-contiguous: rank r evaluates u[x0_r:x1_r]; rank 0 gathers.
    Returns the numpy grid on rank 0 and None elsewhere.  The only collective is the final gather."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    per = (resolution + world - 1) // world
    x0, x1 = slab_bounds(resolution, rank, world)
    dev = next(sdf_network.parameters()).device
    slab = torch.zeros(per, resolution, resolution, dtype=torch.float32, device=dev)
    if x1 > x0:
        sdf_slab(sdf_network, bound_min, bound_max, resolution, x0, x1, out=slab[:x1 - x0])
    parts = [torch.empty_like(slab) for _ in range(world)] if rank == 0 else None
    dist.gather(slab, parts, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat(parts, 0)[:resolution].cpu().numpy()


@torch.no_grad()
def extract_fields_callable(bound_min, bound_max, resolution, query_func):
    """Reference semantics for an arbitrary query_func: 64^3 blocks (models/renderer.py:10-25)."""
    N = 64
    dev = bound_min.device if torch.is_tensor(bound_min) else None
    X = torch.linspace(float(bound_min[0]), float(bound_max[0]), resolution, device=dev).split(N)
    Y = torch.linspace(float(bound_min[1]), float(bound_max[1]), resolution, device=dev).split(N)
    Z = torch.linspace(float(bound_min[2]), float(bound_max[2]), resolution, device=dev).split(N)
    u = np.zeros([resolution, resolution, resolution], dtype=np.float32)
    for xi, xs in enumerate(X):
        for yi, ys in enumerate(Y):
            for zi, zs in enumerate(Z):
                xx, yy, zz = torch.meshgrid(xs, ys, zs, indexing="ij")
                pts = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], -1)
                val = query_func(pts).reshape(len(xs), len(ys), len(zs)).detach().cpu().numpy()
                u[xi * N: xi * N + len(xs), yi * N: yi * N + len(ys), zi * N: zi * N + len(zs)] = val
    return u


def marching_cubes(u, threshold):
    """Host marching cubes.  PyMCubes (the reference's dependency, README.md:36) when installed."""
    try:
        import mcubes
    except ImportError as e:
        raise RuntimeError("rnb_b200: PyMCubes is not installed; the SDF lattice `u` (extract_fields) is the boundary "
                           "of the B200 path -- install PyMCubes for the host marching-cubes step") from e
    return mcubes.marching_cubes(u, threshold)
