"""The loop body of Runner.train_rnb (reference exp_runner.py:168-263) for a scene held on the device.

Same sequence per iteration as the reference -- reseed, pick a view, draw pixels, gather the batch, warm-up or regular
render, RNb loss, zero_grad / backward / step, learning-rate schedule -- with the three optional replacements of
INTEGRATION.md switched on: `DeviceRayBatcher` (8f-1) instead of the CPU fancy-indexing, `FlatAdam` (8f-2) instead of
torch.optim.Adam, and no per-step host synchronisation (the reference's six TensorBoard scalars per step,
exp_runner.py:269-274, are returned as device tensors every `report_freq` steps instead).
Orchestration only: every FLOP happens in the kernels behind `NeuSRenderer`.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .optim import FlatAdam
from .raygen import DeviceRayBatcher
from .scene import learning_rate_factor


def rnb_loss(render_out, true_rgb, mask, n_lights, igr_weight, mask_weight):
    """exp_runner.py:241-256"""
    mask_sum = mask.sum() + 1e-5
    err = ((render_out["color_fine"] - true_rgb) * mask[None, :, :]).reshape(-1, 3)
    color = F.l1_loss(err, torch.zeros_like(err), reduction="sum") / (mask_sum * n_lights)
    mask_loss = F.binary_cross_entropy(render_out["weight_sum"].clip(1e-3, 1.0 - 1e-3), mask)
    return color + render_out["gradient_error"] * igr_weight + mask_loss * mask_weight, color


def train_rnb(renderer, networks, scene, n_iters, batch_size=512, learning_rate=5e-4, learning_rate_alpha=0.05,
              warm_up_end=0, end_iter=None, warm_up_iter=0, anneal_end=0, igr_weight=0.1, mask_weight=0.1,
              no_albedo=False, report_freq=50, log=None, on_iter=None, use_graph=False):
    """networks: the modules whose parameters train (exp_runner.py:105-112 order: nerf, sdf, deviation, colour).
    scene: dict from scene.sphere_scene (or the same tensors of a real Dataset).  Returns (optimizer, history) with
    history = [(iter, loss, color_loss)] sampled every report_freq iterations.
    use_graph: replay one CUDA graph per step (graph_step.GraphedTrainStep; one graph for the warm-up mode, one for the
    regular mode) -- pays in the launch-bound regime of the reference's 512-ray batches; needs a constant
    cos_anneal_ratio (anneal_end = 0, as in the wmask confs)."""
    params = [p for m in networks for p in m.parameters()]
    dev = params[0].device
    opt = FlatAdam(params, lr=learning_rate)
    rb = DeviceRayBatcher(scene["images"], scene["images_warmup"], scene["masks"], scene["light_directions"],
                          scene["intrinsics_all_inv"], scene["pose_all"], device=dev)
    lw_all = scene["light_directions_warmup"].to(dev)
    n_images, n_lights = rb.n_images, rb.n_lights
    end_iter = end_iter if end_iter is not None else n_iters
    g = torch.Generator().manual_seed(0)
    image_perm = torch.randperm(n_images, generator=g)
    history, graphs = [], {}
    for it in range(n_iters):
        if on_iter:
            on_iter(it)
        for grp in opt.param_groups:        # Runner.update_learning_rate runs every iteration (exp_runner.py:320-332)
            grp["lr"] = learning_rate * learning_rate_factor(it, warm_up_end, end_iter, learning_rate_alpha)
        torch.random.manual_seed(it)                                           # exp_runner.py:170
        if it > 0 and it % n_images == 0:
            image_perm = torch.randperm(n_images, generator=g)                 # exp_runner.py:304-306: a new permutation per epoch
        cbn = int(image_perm[it % n_images])
        px = torch.randint(low=0, high=rb.W, size=[batch_size], device="cpu")  # models/dataset.py:356-357
        py = torch.randint(low=0, high=rb.H, size=[batch_size], device="cpu")
        warm = it < warm_up_iter
        b = rb.gather(cbn, px, py, want_lights=not warm)
        mask = (b["mask"] > 0.5).float() if mask_weight > 0.0 else torch.ones_like(b["mask"])
        ratio = 1.0 if anneal_end == 0 else min(1.0, it / anneal_end)          # exp_runner.py:313-317
        if use_graph:
            if anneal_end != 0:
                raise ValueError("use_graph bakes cos_anneal_ratio into the graph: it needs anneal_end = 0")
            if warm not in graphs:
                from .graph_step import GraphedTrainStep
                ex = dict(b, lights_dir=lw_all[cbn].reshape(n_lights, 1, 1, 3) if warm else b["lights_dir"],
                          true_rgb=b["images_warmup"] if warm else b["images"], mask=mask)
                graphs[warm] = GraphedTrainStep(
                    renderer, params, lambda o, rgb, m: rnb_loss(o, rgb, m, n_lights, igr_weight, mask_weight)[0], ex,
                    warmup=warm, no_albedo=no_albedo, cos_anneal_ratio=1.0, reducer=opt.reducer)
            b["lights_dir"] = lw_all[cbn].reshape(n_lights, 1, 1, 3) if warm else b["lights_dir"]
            b["true_rgb"], b["mask"] = (b["images_warmup"] if warm else b["images"]), mask
            loss = color = graphs[warm](b)
            opt.step()
        elif warm:
            true_rgb = b["images_warmup"]
            out = renderer.render_rnb_warmup(b["rays_o"], b["rays_d"], b["near"], b["far"],
                                             lw_all[cbn].reshape(n_lights, 1, 1, 3), cos_anneal_ratio=ratio,
                                             no_albedo=no_albedo)
        else:
            true_rgb = b["images"]
            out = renderer.render_rnb(b["rays_o"], b["rays_d"], b["near"], b["far"], b["lights_dir"],
                                      cos_anneal_ratio=ratio, no_albedo=no_albedo)
        if not use_graph:
            loss, color = rnb_loss(out, true_rgb, mask, n_lights, igr_weight, mask_weight)
            opt.zero_grad()
            loss.backward()
            opt.step()
        if report_freq and (it % report_freq == 0 or it == n_iters - 1):
            history.append((it, float(loss.detach()), float(color.detach())))
            if log:
                log(f"iter {it:6d}  loss {history[-1][1]:.5f}  color {history[-1][2]:.5f}  lr {opt.param_groups[0]['lr']:.2e}")
    return opt, history
