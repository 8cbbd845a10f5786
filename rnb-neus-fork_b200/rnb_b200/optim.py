"""Step epilogue of train_rnb on flat buffers.

The reference builds `torch.optim.Adam(params_to_train, lr=...)` over the 61 parameter tensors of the four networks
(exp_runner.py:105-115), calls `optimizer.zero_grad()` / `loss.backward()` / `optimizer.step()` every iteration
(:259-263), rewrites `param_groups[i]['lr']` from its schedule (:320-332) and stores `optimizer.state_dict()` in the
checkpoint (:380, restored at :368).  `FlatAdam` keeps exactly that interface and checkpoint format, but parameters,
gradients, exp_avg and exp_avg_sq are four flat fp32 buffers -- the gradient buffer is the one the data-parallel
all-reduce works on (`parallel.FlatGradAllReducer`) -- and one launch of `rnb_adam_step` per parameter group replaces
~250 small kernels.  The 1/world scale of the gradient all-reduce can be folded into the same launch (`grad_scale`).

Differences from torch.optim.Adam, all on purpose:
  * every parameter always has a gradient (a view of the flat buffer, zero if nothing flowed into it), so parameters the
    reference would skip because `.grad is None` (colour net under --no_albedo, NeRF with n_outside = 0) get a zero
    update instead of no update, and all parameters share one step count;
  * amsgrad / weight_decay / maximize are not implemented (the reference uses none of them) and raise.
There is no CPU fallback: `step()` needs the parameters on a CUDA device.
"""
from __future__ import annotations

import torch

from . import lib as L
from .parallel import FlatGradAllReducer


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, *,
                 reducer: FlatGradAllReducer = None, grad_scale: float = 1.0):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("rnb_b200.FlatAdam implements the reference's plain Adam (exp_runner.py:115): "
                                      "weight_decay = 0, amsgrad = False")
        defaults = dict(torch.optim.Adam([torch.zeros(1)]).defaults)      # same group keys as a torch Adam checkpoint
        defaults.update(lr=lr, betas=tuple(betas), eps=eps, weight_decay=0, amsgrad=False)
        super().__init__(params, defaults)
        plist = [p for g in self.param_groups for p in g["params"]]
        for p in plist:
            if p.dtype != torch.float32 or not p.requires_grad:
                raise RuntimeError("rnb_b200.FlatAdam: parameters must be trainable float32 tensors")
        if reducer is None:
            reducer = FlatGradAllReducer(plist)
        elif len(reducer.params) != len(plist) or any(a is not b for a, b in zip(reducer.params, plist)):
            raise RuntimeError("rnb_b200.FlatAdam: the reducer must hold the same parameters in the same order")
        self.reducer = reducer
        self.grad_scale = float(grad_scale)
        g = reducer.flat
        self.flat_param = torch.zeros_like(g)
        self.flat_m = torch.zeros_like(g)
        self.flat_v = torch.zeros_like(g)
        self._step = 0
        self._step_t = torch.tensor(0.0)
        # group ranges [start, end) in the flat buffers (groups are contiguous because the reducer keeps parameter order)
        self._ranges = []
        i = 0
        for grp in self.param_groups:
            n = len(grp["params"])
            lo = reducer.offsets[i] if n else 0
            hi = (reducer.offsets[i + n] if i + n < len(plist) else g.numel()) if n else 0
            self._ranges.append((lo, hi))
            i += n
        with torch.no_grad():
            for p, o in zip(plist, reducer.offsets):
                v = self.flat_param[o:o + p.numel()].view_as(p)
                v.copy_(p)
                p.data = v
        self._point_state()

    def _point_state(self):
        for p, o in zip(self.reducer.params, self.reducer.offsets):
            n = p.numel()
            self.state[p] = {"step": self._step_t, "exp_avg": self.flat_m[o:o + n].view_as(p),
                             "exp_avg_sq": self.flat_v[o:o + n].view_as(p)}

    def zero_grad(self, set_to_none: bool = True):
        """Zeroes the flat gradient buffer; `.grad` stays a view of it (never None), whatever `set_to_none` says."""
        self.reducer.zero()

    def _collect_grads(self):
        # gradients arrive as separate tensors (reducer.zero() unsets .grad so that autograd launches nothing per parameter)
        # or were assigned by the caller: one multi-tensor copy folds them into the flat buffer
        self.reducer.collect()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L.require_cuda(self.flat_param, "FlatAdam.step")
        self._collect_grads()
        self._step += 1
        lib = L.load()
        st = L.stream_ptr()
        for (lo, hi), grp in zip(self._ranges, self.param_groups):
            if hi <= lo:
                continue
            if grp.get("weight_decay", 0) != 0 or grp.get("amsgrad", False) or grp.get("maximize", False):
                raise NotImplementedError("rnb_b200.FlatAdam: weight_decay / amsgrad / maximize are not implemented")
            b1, b2 = grp["betas"]
            L.check(lib.rnb_adam_step(L.ptr(self.flat_param[lo:hi]), L.ptr(self.reducer.flat[lo:hi]),
                                      L.ptr(self.flat_m[lo:hi]), L.ptr(self.flat_v[lo:hi]), hi - lo, float(grp["lr"]),
                                      float(b1), float(b2), float(grp["eps"]), self._step, self.grad_scale, st),
                    "adam_step")
        self._step_t.fill_(float(self._step))
        # the kernel wrote the parameters through raw pointers: bump their version counters like an in-place torch op
        # would, so that everything keyed on them (the cached packed weights of ops.py, autograd's saved-tensor checks)
        # sees the update
        torch.autograd.graph.increment_version(self.reducer.params)
        return loss

    def load_state_dict(self, state_dict):
        """Accepts checkpoints written by torch.optim.Adam (exp_runner.py:368) or by this class."""
        super().load_state_dict(state_dict)
        step = 0
        with torch.no_grad():
            self.flat_m.zero_()
            self.flat_v.zero_()
            for p, o in zip(self.reducer.params, self.reducer.offsets):
                s = self.state.get(p, {})
                if "exp_avg" in s:
                    self.flat_m[o:o + p.numel()].view_as(p).copy_(s["exp_avg"])
                    self.flat_v[o:o + p.numel()].view_as(p).copy_(s["exp_avg_sq"])
                    step = max(step, int(float(s["step"])))
        self._step = step
        self._step_t = torch.tensor(float(step))
        self._point_state()
