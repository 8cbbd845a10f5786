import sys, time; sys.path[:0]=['.','tests','rnb-neus-fork_b200']
import numpy as np, torch
from test_gpu_e2e import make_renderer, loss_fn
from rnb_b200 import synth, grid, kernels as K, ops
torch.backends.cuda.matmul.allow_tf32 = False
renderer, sdf, var, col = make_renderer(True)
def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
# grid
bmin, bmax = torch.tensor([-1.01]*3), torch.tensor([1.01]*3)
for R in (128, 256):
    out = torch.empty(R,R,R, device="cuda")
    ms = timeit(lambda: grid.sdf_slab(sdf, bmin, bmax, R, 0, R, out=out))
    n = R**3
    print(f"grid R={R}: {ms:.3f} ms  {n/ms*1e3:.3e} q/s  {n*918016/ms*1e3/1e12:.1f} TFLOP/s")
for B in (512, 8192):
    b = {k: v.cuda() for k, v in synth.make_batch(B, 3, True, 1).items()}
    params = [p for m in (sdf, var) for p in m.parameters()]
    def step():
        for p in params: p.grad = None
        out = renderer.render_rnb_warmup(b["rays_o"], b["rays_d"], b["near"], b["far"], b["lights_dir"], cos_anneal_ratio=1.0, no_albedo=True)
        loss = loss_fn(out, b["true_rgb"], b["mask"], 0.1)
        loss.backward()
    ms = timeit(step)
    print(f"train no_albedo B={B}: {ms:.3f} ms/step  {B/ms*1e3:.3e} rays/s  {B*855433216/ms*1e3/1e12:.1f} TFLOP/s algorithmic")
    # pieces
    pk = ops.packed_sdf_nograd(sdf)
    z, mid = ops.hierarchical_sample(sdf, b["rays_o"], b["rays_d"], b["near"], b["far"], b["t_rand"], 64, 64, 4)
    ms_s = timeit(lambda: ops.hierarchical_sample(sdf, b["rays_o"], b["rays_d"], b["near"], b["far"], b["t_rand"], 64, 64, 4))
    pts = K.points_rays(b["rays_o"], b["rays_d"], mid)
    st = K.SdfStreams(pts.n_pts, "cuda")
    ms_f = timeit(lambda: K.sdf_fwd_grad(pk, pts, st))
    n = pts.n_pts
    d_sdf = torch.randn(n, device="cuda")*1e-4; d_grad = torch.randn(n,3, device="cuda")*1e-5
    scratch = torch.empty(K.L.load().rnb_sdf_bwd_scratch_bytes(n), dtype=torch.uint8, device="cuda")
    ms_b = timeit(lambda: K.sdf_bwd(pk, pts, st, d_sdf, d_grad, None, scratch))
    print(f"   sampling {ms_s:.3f} ms | sdf_fwd_grad {ms_f:.3f} ms ({n*(1049088+917504)/ms_f*1e3/1e12:.1f} TF/s) | sdf_bwd {ms_b:.3f} ms ({n*3913216/ms_b*1e3/1e12:.1f} TF/s)")
