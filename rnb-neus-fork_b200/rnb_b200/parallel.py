"""Data-parallel training step support: one process per GPU, rays sharded across ranks, ONE all-reduce of the
parameter gradients per step (NCCL over NVLink / NVSwitch; gloo on CPU for the host-logic tests).

The reference is single-process; this is what the north_star adds.  Parameters are replicated, every
rank renders its own rays, `.grad` of all parameters are views into one flat fp32 buffer so the collective is a single
all-reduce of 2.1-2.7 MB (NCCL `ReduceOp.AVG`: the 1/world scale happens inside the collective).  Loss normalisers (mask_sum, the eikonal
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

    def all_reduce_sum(self):
        """Gather into the flat buffer and SUM over the ranks (ExactBatch: every rank's loss is its share of one global-
        batch loss, so the shares add)."""
        self.collect()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        return self.flat

    def all_reduce(self):
        """Gather into the flat buffer and average over the ranks: ONE collective per step and nothing else -- NCCL
        averages inside the all-reduce (ReduceOp.AVG); gloo (CPU tests) has no AVG, there it is a sum and a scale."""
        self.collect()
        if not (dist.is_available() and dist.is_initialized()):
            return self.flat
        world = dist.get_world_size(self.group)
        if world > 1:
            if self.flat.is_cuda and dist.get_backend(self.group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
            else:
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


class ExactBatch:
    """Optional exact global-batch equivalence ("or all-reduce the two scalars first").

    With plain DDP semantics each rank normalises its loss by its own batch: the colour term by its own mask_sum
    (exp_runner.py:194, 247), the eikonal term by its own count of points inside the relaxed sphere
    (models/renderer.py:540) -- N ranks x B rays then differ from one N*B-ray batch by O(1/B).  ExactBatch removes that:

        eb = ExactBatch(renderer)                    # renderer.dp_exact_group = the group: eikonal num / den summed over ranks
        out = renderer.render_rnb(...)               # out['gradient_error'] is the GLOBAL mean on every rank
        loss = eb.loss(out, true_rgb, mask, igr_weight, mask_weight)      # this rank's share of the global-batch loss
        loss.backward(); reducer.all_reduce_sum()    # shares add: SUM, not mean

    `loss` = colour L1 sum of this rank / (global mask_sum * L) + igr * (global eikonal mean, differentiated through this
    rank's points) + mask_weight * local BCE sum / (global ray count); its gradients summed over the ranks equal the gradient
    of the reference loss on the concatenated batch (checked under NCCL with the real kernels by bench.py's dp_check)."""

    def __init__(self, renderer, group=None):
        self.group = group
        self.renderer = renderer
        renderer.dp_exact_group = group if group is not None else True

    def world(self):
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    def loss(self, out, true_rgb, mask, igr_weight=0.1, mask_weight=0.1):
        import torch.nn.functional as F
        world = self.world()
        norm = torch.stack([mask.sum(), torch.tensor(float(mask.shape[0]), device=mask.device)])
        if world > 1:
            dist.all_reduce(norm, op=dist.ReduceOp.SUM, group=self.group)
        mask_sum, n_rays = norm[0] + 1e-5, norm[1]
        err = ((out["color_fine"] - true_rgb) * mask[None, :, :]).reshape(-1, 3)
        color = F.l1_loss(err, torch.zeros_like(err), reduction="sum") / (mask_sum * true_rgb.shape[0])
        bce = F.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), mask, reduction="sum") / n_rays
        # gradient_error is the global ratio on every rank and its backward is this rank's share of it; its VALUE would be
        # counted `world` times by the sum over ranks, which does not matter for the gradients (and the logged loss is
        # reported by total()).
        return color + out["gradient_error"] * igr_weight + bce * mask_weight

    def total(self, loss, out, igr_weight=0.1):
        """The global-batch loss value from the per-rank shares (for logging): sum of shares minus the (world - 1) extra
        copies of the eikonal term."""
        world = self.world()
        t = loss.detach().clone()
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            t = t - (world - 1) * igr_weight * out["gradient_error"].detach()
        return t
