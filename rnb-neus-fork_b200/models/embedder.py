"""Positional encoding with the reference's interface (reference models/embedder.py:58-74).

`get_embedder(multires, input_dims)` -> (embed_fn, out_dim); output column order
x | sin(2^0 x) | cos(2^0 x) | sin(2^1 x) | ... (reference models/embedder.py:21-45).
The hot path never calls this: the CUDA kernels generate the encoding in registers
(csrc/pe.cuh).  It exists so code that imports `models.embedder` keeps working.
"""
import torch


class Embedder:
    def __init__(self, input_dims, num_freqs, include_input=True):
        self.input_dims = input_dims
        self.num_freqs = num_freqs
        self.include_input = include_input
        self.out_dim = input_dims * ((1 if include_input else 0) + 2 * num_freqs)

    def embed(self, inputs):
        freqs = 2.0 ** torch.arange(self.num_freqs, dtype=inputs.dtype, device=inputs.device)
        ang = inputs.unsqueeze(-2) * freqs.view(-1, 1)                   # [..., L, d]
        sc = torch.stack([torch.sin(ang), torch.cos(ang)], dim=-2)      # [..., L, 2, d]
        sc = sc.flatten(-3)
        return torch.cat([inputs, sc], -1) if self.include_input else sc


def get_embedder(multires, input_dims=3):
    obj = Embedder(input_dims, multires)
    return obj.embed, obj.out_dim
