"""Data-parallel training step support: one process per GPU, rays sharded across ranks, ONE all-reduce of the
parameter gradients per step (NCCL over NVLink / NVSwitch; gloo on CPU for the host-logic tests).

The reference is single-process; this is what the north_star adds.  Parameters are replicated, every
rank renders its own rays, `.grad` of all parameters are views into one flat fp32 buffer so the collective is a single
`all_reduce(sum)` of 2.1-2.7 MB followed by a scale by 1/world.  Loss normalisers (mask_sum, the eikonal
denominator) stay per-rank, i.e. DDP "mean of per-rank means" semantics.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class FlatGradAllReducer:
    def __init__(self, params, group=None, align=4):
        """align: every parameter's slice starts at a multiple of `align` floats (16 bytes by default, so slices keep
        the alignment vector loads need); the padding stays zero."""
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.offsets = []
        o = 0
        for p in self.params:
            self.offsets.append(o)
            o += -(-p.numel() // align) * align
        dev = self.params[0].device
        self.flat = torch.zeros(o, dtype=torch.float32, device=dev)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for p, o in zip(self.params, self.offsets)]
        self.attach()

    def attach(self):
        """Point every .grad at its slice of the flat buffer (autograd then accumulates in place)."""
        for p, v in zip(self.params, self.views):
            p.grad = v

    def zero(self):
        self.flat.zero_()
        self.attach()

    def all_reduce(self):
        """Sum over ranks, divide by the world size.  One collective per step."""
        if not (dist.is_available() and dist.is_initialized()):
            return self.flat
        world = dist.get_world_size(self.group)
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / world)
        return self.flat

    @property
    def nbytes(self):
        return self.flat.numel() * 4


def rank_seed(iter_i: int, rank: int, world: int) -> int:
    """The reference reseeds with manual_seed(iter_i) every iteration (exp_runner.py:170); under data parallelism
    every rank must draw different pixels, so the seed is offset by the rank."""
    return iter_i * world + rank
