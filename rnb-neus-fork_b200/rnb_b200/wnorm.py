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


class _FoldAll(torch.autograd.Function):
    @staticmethod
    def forward(ctx, n, *vg):
        ctx.set_materialize_grads(False)
        vs = [t.detach() for t in vg[:n]]
        gs = [t.detach() for t in vg[n:]]
        for t in vs + gs:
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("rnb_b200.wnorm: weight_v / weight_g must be contiguous float32 tensors")
        eo, ro, ne, nr = _plan(vs)
        dev = vs[0].device
        wbuf = torch.empty(ne, dtype=torch.float32, device=dev)
        nbuf = torch.empty(nr, dtype=torch.float32, device=dev)
        arr = (L.WnLayer * n)()
        wp, npn = wbuf.data_ptr(), nbuf.data_ptr()
        for i in range(n):
            a = arr[i]
            a.rows, a.cols = vs[i].shape
            a.v, a.g = vs[i].data_ptr(), gs[i].data_ptr()
            a.w, a.norm = wp + 4 * eo[i], npn + 4 * ro[i]
        L.check(L.load().rnb_weight_norm_fold(arr, n, L.stream_ptr()), "weight_norm_fold")
        ctx.n, ctx.vs, ctx.gs, ctx.nbuf, ctx.plan = n, vs, gs, nbuf, (eo, ro, ne, nr)
        return tuple(wbuf[eo[i]:eo[i] + vs[i].numel()].view(vs[i].shape) for i in range(n))

    @staticmethod
    def backward(ctx, *dws):
        n, vs, gs = ctx.n, ctx.vs, ctx.gs
        eo, ro, ne, nr = ctx.plan
        idx = [i for i in range(n) if dws[i] is not None]
        dvs, dgs = [None] * n, [None] * n
        if idx:
            dev = vs[0].device
            dvbuf = torch.empty(ne, dtype=torch.float32, device=dev)
            dgbuf = torch.empty(nr, dtype=torch.float32, device=dev)
            keep = []
            arr = (L.WnLayer * len(idx))()
            npn, dvp, dgp = ctx.nbuf.data_ptr(), dvbuf.data_ptr(), dgbuf.data_ptr()
            for k, i in enumerate(idx):
                dw = dws[i]
                if dw.dtype != torch.float32 or not dw.is_contiguous():
                    dw = dw.float().contiguous()
                keep.append(dw)
                a = arr[k]
                a.rows, a.cols = vs[i].shape
                a.v, a.g, a.w = vs[i].data_ptr(), gs[i].data_ptr(), dw.data_ptr()
                a.norm, a.dv, a.dg = npn + 4 * ro[i], dvp + 4 * eo[i], dgp + 4 * ro[i]
                dvs[i] = dvbuf[eo[i]:eo[i] + vs[i].numel()].view(vs[i].shape)
                dgs[i] = dgbuf[ro[i]:ro[i] + vs[i].shape[0]].view(gs[i].shape)
            L.check(L.load().rnb_weight_norm_vjp(arr, len(idx), L.stream_ptr()), "weight_norm_vjp")
        return (None, *dvs, *dgs)


def fold_all(vs, gs):
    """vs: weight_v [out,in] per layer, gs: weight_g [out,1] per layer -> list of effective weights (autograd-connected)."""
    L.require_cuda(vs[0], "weight_norm")
    return list(_FoldAll.apply(len(vs), *vs, *gs))
