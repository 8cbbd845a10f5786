"""SDFNetwork / RenderingNetwork / NeRF / SingleVarianceNetwork with the reference's
constructor and forward signatures and state_dict keys (reference models/fields.py:8-325),
evaluated by the sm_100a kernels in rnb_b200 (tcgen05 MLP chains; csrc/chain.cuh).

Parameters are created with the same torch RNG call sequence as the reference, so
`torch.manual_seed(s)` gives bit-identical initial weights and reference checkpoints
(`linK.weight_g / weight_v / bias`, `pts_linears.N.weight`, `variance`;
reference exp_runner.py:355-386) load unchanged in both directions.
"""
import os as _os

import numpy as np
import torch
import torch.nn as nn

from models.embedder import get_embedder


def _ops():
    from rnb_b200 import ops
    return ops


def _warn_forward_only(module, what):
    """The stand-alone RenderingNetwork.forward / NeRF.forward / render() with n_outside > 0 are forward-only here (their
    reference callers, validate_mesh_texture and render_novel_image, detach the results).  The reference modules are
    differentiable: say so loudly instead of silently returning a detached tensor to a caller that trains through them."""
    if torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters()):
        import warnings
        warnings.warn(f"rnb_b200: {what} is forward-only (no autograd graph is recorded); its result is detached. "
                      "Train through NeuSRenderer.render_rnb / render_rnb_warmup, or call it under torch.no_grad().",
                      RuntimeWarning, stacklevel=3)


def _use_fused_weight_norm():
    """All layers of a network are folded by ONE launch (rnb_weight_norm_fold) and their VJP is one more
    (rnb_b200/wnorm.py) instead of torch's 24 per-layer kernels per step -- in eager steps and under CUDA-graph capture
    alike, so both modes run the same arithmetic.  Round 1 used it only under capture because the node's Python side cost
    1 ms of host time per 512-ray step; with the layer tables cached it costs what torch's C++ per-layer ops cost
    (b512_noalbedo 1.97 ms either way, profiles/r02_notes.md).  RNB_FUSED_WN=0 restores torch._weight_norm."""
    force = _os.environ.get("RNB_FUSED_WN")
    if force is not None:
        return force not in ("0", "")
    return True


class _WeightNormMLP(nn.Module):
    """Shared plumbing: layers are attributes lin0..linN-1 (legacy weight_norm => weight_g/weight_v)."""

    def effective_weights(self):
        """[(W_l fp32 [out,in], b_l)] with weight-norm folded (reference models/fields.py:72-74); autograd turns the
        kernels' dW into (dg, dv).  All layers are folded by ONE launch (rnb_b200/wnorm.py) and their VJP is one more;
        CPU modules (host-side contract tests) and RNB_FUSED_WN=0 use torch's per-layer _weight_norm."""
        lins = [getattr(self, "lin" + str(l)) for l in range(self.num_layers - 1)]
        if all(hasattr(lin, "weight_g") for lin in lins) and lins[0].weight_v.is_cuda and _use_fused_weight_norm():
            from rnb_b200 import wnorm
            Ws = wnorm.fold_all([lin.weight_v for lin in lins], [lin.weight_g for lin in lins])
            return [(W, lin.bias) for W, lin in zip(Ws, lins)]
        out = []
        for lin in lins:
            if hasattr(lin, "weight_g"):
                W = torch._weight_norm(lin.weight_v, lin.weight_g, 0)
            else:
                W = lin.weight
            out.append((W, lin.bias))
        return out


class SDFNetwork(_WeightNormMLP):
    """reference models/fields.py:8-127"""

    def __init__(self, d_in, d_out, d_hidden, n_layers, skip_in=(4,), multires=0, bias=0.5, scale=1,
                 geometric_init=True, weight_norm=True, inside_outside=False):
        super().__init__()
        dims = [d_in] + [d_hidden] * n_layers + [d_out]
        self.embed_fn_fine = None
        if multires > 0:
            self.embed_fn_fine, dims[0] = get_embedder(multires, input_dims=d_in)
        self.multires = multires
        self.num_layers = len(dims)
        self.skip_in = tuple(skip_in)
        self.scale = scale
        self.dims = dims
        last = self.num_layers - 2
        for l in range(self.num_layers - 1):
            n_out = dims[l + 1] - dims[0] if (l + 1) in self.skip_in else dims[l + 1]
            lin = nn.Linear(dims[l], n_out)
            if geometric_init:
                with torch.no_grad():
                    if l == last:
                        sign = -1.0 if inside_outside else 1.0
                        nn.init.normal_(lin.weight, mean=sign * np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                        nn.init.constant_(lin.bias, -sign * bias)
                    elif multires > 0 and l == 0:
                        nn.init.constant_(lin.bias, 0.0)
                        nn.init.constant_(lin.weight[:, 3:], 0.0)
                        nn.init.normal_(lin.weight[:, :3], 0.0, np.sqrt(2) / np.sqrt(n_out))
                    elif multires > 0 and l in self.skip_in:
                        nn.init.constant_(lin.bias, 0.0)
                        nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(n_out))
                        nn.init.constant_(lin.weight[:, -(dims[0] - 3):], 0.0)
                    else:
                        nn.init.constant_(lin.bias, 0.0)
                        nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(n_out))
            if weight_norm:
                lin = nn.utils.weight_norm(lin)
            setattr(self, "lin" + str(l), lin)
        self.activation = nn.Softplus(beta=100)

    # -- hot-path entry points (all CUDA) ---------------------------------
    def forward(self, inputs):
        """[N,3] -> [N,d_out] (col 0 sdf/scale, cols 1: features); differentiable w.r.t. parameters."""
        return _ops().sdf_forward(self, inputs, want_grad=False)[0]

    def sdf(self, x):
        ops = _ops()
        if not torch.is_grad_enabled() or not any(p.requires_grad for p in self.parameters()):
            return ops.sdf_only(self, x)
        return self.forward(x)[:, :1]

    def sdf_hidden_appearance(self, x):
        return self.forward(x)

    def gradient(self, x):
        """d sdf / d x, [N,1,3]; analytic dx-chain kernel instead of autograd (reference :114-127)."""
        return _ops().sdf_forward(self, x, want_grad=True)[1].unsqueeze(1)

    def forward_with_gradient(self, x):
        """(out [N,d_out], grad [N,3]) from one fused pass -- what render_core_mvps needs."""
        return _ops().sdf_forward(self, x, want_grad=True)


class RenderingNetwork(_WeightNormMLP):
    """reference models/fields.py:131-215"""

    def __init__(self, d_feature, mode, d_in, d_out, d_hidden, n_layers, weight_norm=True, multires_view=0,
                 squeeze_out=True):
        super().__init__()
        self.mode = mode
        self.squeeze_out = squeeze_out
        dims = [d_in + d_feature] + [d_hidden] * n_layers + [d_out]
        self.embedview_fn = None
        self.multires_view = multires_view
        if multires_view > 0:
            self.embedview_fn, input_ch = get_embedder(multires_view)
            if mode == "no_view_dir":
                dims[0] += 2 * (input_ch - 3)
            if mode == "ps":
                dims[0] = input_ch
        self.num_layers = len(dims)
        self.dims = dims
        for l in range(self.num_layers - 1):
            lin = nn.Linear(dims[l], dims[l + 1])
            if weight_norm:
                lin = nn.utils.weight_norm(lin)
            setattr(self, "lin" + str(l), lin)
        self.relu = nn.ReLU()

    def forward(self, points, normals, view_dirs, feature_vectors):
        _warn_forward_only(self, "RenderingNetwork.forward called on its own")
        return _ops().albedo_forward(self, points, normals, view_dirs, feature_vectors)


class NeRF(nn.Module):
    """NeRF++ background field; reference models/fields.py:219-314."""

    def __init__(self, D=8, W=256, d_in=3, d_in_view=3, multires=0, multires_view=0, output_ch=4, skips=[4],
                 use_viewdirs=False):
        super().__init__()
        self.D, self.W, self.d_in, self.d_in_view = D, W, d_in, d_in_view
        self.input_ch, self.input_ch_view = 3, 3
        self.embed_fn = self.embed_fn_view = None
        self.multires, self.multires_view = multires, multires_view
        if multires > 0:
            self.embed_fn, self.input_ch = get_embedder(multires, input_dims=d_in)
        if multires_view > 0:
            self.embed_fn_view, self.input_ch_view = get_embedder(multires_view, input_dims=d_in_view)
        self.skips = skips
        self.use_viewdirs = use_viewdirs
        layers = [nn.Linear(self.input_ch, W)]
        for i in range(D - 1):
            layers.append(nn.Linear(W + self.input_ch, W) if i in self.skips else nn.Linear(W, W))
        self.pts_linears = nn.ModuleList(layers)
        self.views_linears = nn.ModuleList([nn.Linear(self.input_ch_view + W, W // 2)])
        if use_viewdirs:
            self.feature_linear = nn.Linear(W, W)
            self.alpha_linear = nn.Linear(W, 1)
            self.rgb_linear = nn.Linear(W // 2, 3)
        else:
            self.output_linear = nn.Linear(W, output_ch)

    def forward(self, input_pts, input_views):
        assert self.use_viewdirs, "only the use_viewdirs head exists (reference models/fields.py:313-314)"
        _warn_forward_only(self, "NeRF.forward")
        return _ops().nerf_forward(self, input_pts, input_views)


class SingleVarianceNetwork(nn.Module):
    """reference models/fields.py:317-325"""

    def __init__(self, init_val):
        super().__init__()
        self.register_parameter("variance", nn.Parameter(torch.tensor(init_val)))

    def forward(self, x):
        return torch.ones([len(x), 1], device=self.variance.device) * torch.exp(self.variance * 10.0)
