"""Weight-norm fold of every layer of a network as ONE autograd node (two launches per step instead of 24).

Reference: each layer is `nn.utils.weight_norm(lin)` (models/fields.py:72-74, 161-162); torch folds and back-propagates
layer by layer.  `fold_all(vs, gs)` returns the effective weights W_l = g_l v_l / |v_l|_row of all layers from one
`rnb_weight_norm_fold` launch; its backward turns the kernels' dW_l into (dv_l, dg_l) with one `rnb_weight_norm_vjp`
launch.  Outputs of one call are views of one buffer (two allocations per direction, not two per layer).
CUDA only, like the rest of the path: CPU modules keep torch's own `_weight_norm` (host-side contract tests).
"""
from __future__ import annotations

import torch

from . import lib as L


def _plan(vs):
    """-> (element offsets of the [rows, cols] blocks, row offsets, totals); blocks start at multiples of 4 floats"""
    eo, ro, e, r = [], [], 0, 0
    for v in vs:
        eo.append(e)
        ro.append(r)
        e += -(-v.numel() // 4) * 4
        r += -(-v.shape[0] // 4) * 4
    return eo, ro, e, r


_PLANS = {}


def _cached_plan(vs, gs):
    """Everything about one set of layers that does not change from step to step: block offsets, and ctypes layer tables
    with the parameter pointers / shapes already filled in (the parameters are views of FlatAdam's flat buffer or plain
    nn.Parameters: their addresses are stable).  Building the tables was most of this node's host time, which is what
    made it slower than torch's per-layer op in eager mode (profiles/r01_notes.md, session 4)."""
    key = tuple((v.data_ptr(), g.data_ptr(), v.shape[0], v.shape[1]) for v, g in zip(vs, gs))
    plan = _PLANS.get(key)
    if plan is None:
        for t in list(vs) + list(gs):
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("rnb_b200.wnorm: weight_v / weight_g must be contiguous float32 tensors")
        n = len(vs)
        eo, ro, ne, nr = _plan(vs)
        fwd, bwd = (L.WnLayer * n)(), (L.WnLayer * n)()
        for i in range(n):
            for a in (fwd[i], bwd[i]):
                a.rows, a.cols = vs[i].shape
                a.v, a.g = vs[i].data_ptr(), gs[i].data_ptr()
        plan = dict(n=n, eo=eo, ro=ro, ne=ne, nr=nr, fwd=fwd, bwd=bwd, shapes=[tuple(v.shape) for v in vs],
                    numel=[v.numel() for v in vs], gshapes=[tuple(g.shape) for g in gs])
        if len(_PLANS) > 64:
            _PLANS.clear()
        _PLANS[key] = plan
    return plan


class _FoldAll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, n, *vg):
        ctx.set_materialize_grads(False)
        vs, gs = vg[:n], vg[n:]
        plan = _cached_plan(vs, gs)
        eo, ro = plan["eo"], plan["ro"]
        dev = vs[0].device
        wbuf = torch.empty(plan["ne"], dtype=torch.float32, device=dev)
        nbuf = torch.empty(plan["nr"], dtype=torch.float32, device=dev)
        arr = plan["fwd"]
        wp, npn = wbuf.data_ptr(), nbuf.data_ptr()
        for i in range(n):
            arr[i].w, arr[i].norm = wp + 4 * eo[i], npn + 4 * ro[i]
        L.check(L.load().rnb_weight_norm_fold(arr, n, L.stream_ptr()), "weight_norm_fold")
        ctx.plan, ctx.nbuf, ctx.dev = plan, nbuf, dev
        return tuple(wbuf[eo[i]:eo[i] + plan["numel"][i]].view(plan["shapes"][i]) for i in range(n))

    @staticmethod
    def backward(ctx, *dws):
        plan = ctx.plan
        n, eo, ro = plan["n"], plan["eo"], plan["ro"]
        dvs, dgs = [None] * n, [None] * n
        if any(d is not None for d in dws):
            dvbuf = torch.empty(plan["ne"], dtype=torch.float32, device=ctx.dev)
            dgbuf = torch.empty(plan["nr"], dtype=torch.float32, device=ctx.dev)
            keep = []
            arr = plan["bwd"]
            npn, dvp, dgp = ctx.nbuf.data_ptr(), dvbuf.data_ptr(), dgbuf.data_ptr()
            k = 0
            for i in range(n):
                dw = dws[i]
                if dw is None:
                    continue
                if dw.dtype != torch.float32 or not dw.is_contiguous():
                    dw = dw.float().contiguous()
                keep.append(dw)
                a = arr[k]
                if k != i:       # compact the table when some layers received no gradient (rows/cols/v/g of layer i)
                    src = plan["fwd"][i]
                    a.rows, a.cols, a.v, a.g = src.rows, src.cols, src.v, src.g
                a.w = dw.data_ptr()
                a.norm, a.dv, a.dg = npn + 4 * ro[i], dvp + 4 * eo[i], dgp + 4 * ro[i]
                dvs[i] = dvbuf[eo[i]:eo[i] + plan["numel"][i]].view(plan["shapes"][i])
                dgs[i] = dgbuf[ro[i]:ro[i] + plan["shapes"][i][0]].view(plan["gshapes"][i])
                k += 1
            L.check(L.load().rnb_weight_norm_vjp(arr, k, L.stream_ptr()), "weight_norm_vjp")
            if k != n:           # the compacted table no longer matches the layer order: rebuild it next time
                for i in range(n):
                    src, a = plan["fwd"][i], arr[i]
                    a.rows, a.cols, a.v, a.g = src.rows, src.cols, src.v, src.g
        return (None, *dvs, *dgs)


def fold_all(vs, gs):
    """vs: weight_v [out,in] per layer, gs: weight_g [out,1] per layer -> list of effective weights (autograd-connected)."""
    L.require_cuda(vs[0], "weight_norm")
    return list(_FoldAll.apply(len(vs), *vs, *gs))
