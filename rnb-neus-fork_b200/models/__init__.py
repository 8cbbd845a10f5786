"""Drop-in replacement for the reference's `models` package on the RNb-NeuS hot path.

Put `rnb-neus-fork_b200/` on PYTHONPATH in place of the reference's own directory and
`from models.fields import ...` / `from models.renderer import NeuSRenderer`
(reference exp_runner.py:13-15) resolve here.  Same classes, constructor kwargs,
method signatures, returned dict keys and state_dict layout; the compute runs in the
sm_100a CUDA library `rnb_b200/librnb_b200.so` (no CPU fallback).
"""
