"""Helpers shared by the GPU parity tests."""
import numpy as np
import torch

from rnb_b200 import synth


def build_nets(perturb=True, device="cuda"):
    from models import fields
    torch.manual_seed(0)
    conf = synth.WMASK_CONF
    nerf = fields.NeRF(**conf["nerf"])
    sdf = fields.SDFNetwork(**conf["sdf_network"])
    var = fields.SingleVarianceNetwork(**conf["variance_network"])
    col = fields.RenderingNetwork(**conf["rendering_network"])
    if perturb:
        synth.perturb_state_dict_(sdf, synth.SDF_NOISE, 5)
        synth.perturb_state_dict_(col, synth.COLOR_NOISE, 6)
        with torch.no_grad():
            var.variance.fill_(synth.TRAINED_VARIANCE)
    return nerf.to(device), sdf.to(device), var.to(device), col.to(device)


def np_state(module):
    return {k: v.detach().double().cpu().numpy() for k, v in module.state_dict().items()}
