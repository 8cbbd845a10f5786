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
        """Point every .grad at its slice of the flat buffer."""
        for p, v in zip(self.params, self.views):
            p.grad = v

    def zero(self):
        """Start of a step: gradients are dropped (`.grad = None`), nothing is launched.  With `.grad` unset autograd hands
        each parameter's gradient over without a kernel; with the views attached it would run one in-place add per
        parameter (42 tiny launches per train_rnb step).  `collect()` then moves them into the flat buffer in one
        multi-tensor copy."""
        for p in self.params:
            p.grad = None
        self._collected = False

    def collect(self):
        """After backward: gather the parameters' gradients into the flat buffer (one foreach copy) and re-attach the
        views, so `.grad` of every parameter is its slice of `flat`.  Parameters that received no gradient read zero.
        Idempotent until the next zero()."""
        if getattr(self, "_collected", False) and all(p.grad is v for p, v in zip(self.params, self.views)):
            return self.flat
        dst, src = [], []
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is v:
                continue
            if g is None:
                v.zero_()
            elif g.data_ptr() != v.data_ptr():
                dst.append(v)
                src.append(g.detach() if g.shape == v.shape else g.detach().reshape(v.shape))
        if dst:
            with torch.no_grad():
                torch._foreach_copy_(dst, src)
        self.attach()
        self._collected = True
        return self.flat

    def all_reduce(self):
        """Gather into the flat buffer, sum over ranks, divide by the world size.  One collective per step."""
        self.collect()
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
