"""One CUDA graph per training step (launch-bound regime).

At the reference's own batch size (512 rays, confs/wmask_rnb.conf:26) a train_rnb step is ~30 kernels of this library
plus ~150 parameter-sized torch kernels (weight-norm forward/backward, loss, autograd glue): 1.2 ms of GPU work spread
over 1.9 ms of CPU launch time.  `GraphedTrainStep` captures

    zero grads -> render_rnb[_warmup] -> loss -> backward

into one CUDA graph (torch.cuda.CUDAGraph; the library's launches go to torch's current stream, so they are captured
like any other kernel) and replays it per step after copying the batch into static input buffers.  The gradient buffer
is the flat one of `parallel.FlatGradAllReducer`, so the NCCL all-reduce and the optimiser step stay outside the graph.

What is baked in at capture time: batch size, warm-up / regular mode, `no_albedo`, `cos_anneal_ratio` (a kernel
argument; wmask confs keep it at 1.0 -- re-capture when it changes), the light-direction layout.  Random numbers are
drawn inside the graph through torch's graph-safe CUDA generator, so every replay gets fresh jitter.
"""
from __future__ import annotations

import torch

from .parallel import FlatGradAllReducer


class GraphedTrainStep:
    def __init__(self, renderer, params, loss_fn, example_batch, warmup=True, no_albedo=False, cos_anneal_ratio=1.0,
                 reducer: FlatGradAllReducer = None, n_warmup=3):
        """example_batch: dict with rays_o, rays_d, near, far, lights_dir, true_rgb, mask (CUDA tensors of the shapes every
        later batch will have).  loss_fn(render_out, true_rgb, mask) -> scalar loss."""
        self.renderer, self.loss_fn = renderer, loss_fn
        self.red = reducer if reducer is not None else FlatGradAllReducer(list(params))
        self.keys = ("rays_o", "rays_d", "near", "far", "lights_dir", "true_rgb", "mask")
        self.static = {k: example_batch[k].detach().clone().contiguous() for k in self.keys}
        self.fn = renderer.render_rnb_warmup if warmup else renderer.render_rnb
        self.kw = dict(cos_anneal_ratio=float(cos_anneal_ratio), no_albedo=bool(no_albedo))
        # warm-up on a side stream (lazy initialisations, allocator pools), as torch.cuda.graphs requires
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(n_warmup):
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        from . import lib as L
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.launch_count()
        with torch.cuda.graph(self.graph):
            self.loss = self._body()
        self.launches_per_replay = L.launch_count() - n0     # kernels of this library inside the graph (per replay)

    def _body(self):
        b = self.static
        self.red.zero()
        out = self.fn(b["rays_o"], b["rays_d"], b["near"], b["far"], b["lights_dir"], **self.kw)
        loss = self.loss_fn(out, b["true_rgb"], b["mask"])
        loss.backward()
        self.red.collect()          # inside the graph: the flat gradient buffer is what a replay leaves behind
        return loss.detach()

    def __call__(self, batch):
        """Copies `batch` into the static buffers (device-to-device or pinned host-to-device, asynchronous), replays the
        graph and returns the (static) loss tensor.  Gradients are in `.grad` of the parameters = views of reducer.flat."""
        for k in self.keys:
            self.static[k].copy_(batch[k], non_blocking=True)
        self.graph.replay()
        return self.loss
