"""Timeline of one CTA of a chain kernel from an RNB_TRACE build (clock64 stamps of CTA 0):
    RNB_OUT=../rnb_b200/librnb_b200_trace.so bash rnb-neus-fork_b200/csrc/build.sh -DRNB_TRACE
    RNB_B200_LIB=$PWD/rnb-neus-fork_b200/rnb_b200/librnb_b200_trace.so python profiles/_trace_chain.py k1|k2|k3a
Roles: 0 producer per slice [wait_empty_start, empty_seen, issued], 1 MMA per slice [wait_full_start, full_seen,
mma_issued], 2 MMA per step [wait_a_start, a_ready_seen, acc_committed], 3 epilogue thread 64 per step
[wait_acc_start, acc_seen, signal]."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "rnb-neus-fork_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rnb_b200 import kernels as K, lib as L, ops, synth  # noqa: E402
from test_gpu_e2e import make_renderer  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "k1"
renderer, sdf, var, col = make_renderer(True)
lib = L.load()
raw = C.CDLL(L.LIB_PATH)
trace = torch.zeros(4 * 4096 * 4, dtype=torch.int64, device="cuda")
B = 8192
b = {k: v.cuda() for k, v in synth.make_batch(B, 3, True, 1).items()}
pk = ops.packed_sdf_nograd(sdf)
z, mid = ops.hierarchical_sample(sdf, b["rays_o"], b["rays_d"], b["near"], b["far"], b["t_rand"], 64, 64, 4)
pts = K.points_rays(b["rays_o"], b["rays_d"], mid)
st = K.SdfStreams(pts.n_pts, "cuda")
n = pts.n_pts
d_sdf = torch.randn(n, device="cuda") * 1e-4
d_grad = torch.randn(n, 3, device="cuda") * 1e-5
scratch = torch.empty(lib.rnb_sdf_bwd_scratch_bytes(n), dtype=torch.uint8, device="cuda")
K.sdf_fwd_grad(pk, pts, st)
from rnb_b200 import albedo as A  # noqa: E402
sdf_out, grad_out = K.sdf_fwd_grad(pk, pts, st)[:2]
with torch.no_grad():
    col_flat = [t.detach().contiguous() for wb in col.effective_weights() for t in wb]
actx = A.forward(col_flat, pts, grad_out.contiguous(), st)
d_alb = torch.randn(n, 3, device="cuda") * 1e-4
fn = {"k1": lambda: K.sdf_fwd(pk, pts), "k2": lambda: K.sdf_fwd_grad(pk, pts, st),
      "k3a": lambda: K.sdf_bwd(pk, pts, st, d_sdf, d_grad, None, scratch),
      "alb_fwd": lambda: A.forward(col_flat, pts, grad_out.contiguous(), st),
      "alb_bwd": lambda: A.backward(actx, d_alb)}[which]
for _ in range(2):
    fn()
torch.cuda.synchronize()
raw.rnb_trace_set.argtypes = [C.c_void_p]
raw.rnb_trace_set(trace.data_ptr())
fn()
torch.cuda.synchronize()
raw.rnb_trace_set(None)
t = trace.cpu().numpy().reshape(4, 4096, 4)
n_steps = {"k1": 8, "k2": 17, "k3a": 16, "alb_fwd": 3, "alb_bwd": 3}[which]
t0 = t[2, 0, 1]
np.save(os.path.join(ROOT, "gpurun_out", f"trace_{which}.npy"), t)
# averages over the tiles of CTA 0 (first two and last skipped), per step index, seen by epilogue thread 64
t = t.astype(np.int64)
n_tiles = int((t[3, :, 1] > 0).sum()) // n_steps
print(f"{which}: CTA 0, {n_tiles} tiles; mean cycles per step")
print("step | epilogue waits for the accumulator (= MMA phase as the epilogue sees it) | epilogue work (acc seen -> signal)")
tot_w = tot_e = 0.0
for st in range(n_steps):
    idx = np.array([tile * n_steps + st for tile in range(2, n_tiles - 1)])
    wait = (t[3, idx, 1] - t[3, idx, 0]).mean()
    work = (t[3, idx + 1, 2] - t[3, idx, 1]).mean() if st + 1 < n_steps else float("nan")
    tot_w += wait
    tot_e += 0.0 if work != work else work
    print(f"{st:4d} | {wait:8.0f} | {work:8.0f}")
span = (t[3, (n_tiles - 2) * n_steps, 1] - t[3, 2 * n_steps, 1]) / (n_tiles - 4)
print(f"sum of waits {tot_w:.0f}, sum of work {tot_e:.0f} (last step's work not recorded), cycles per tile {span:.0f}")
