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
fn = {"k1": lambda: K.sdf_fwd(pk, pts), "k2": lambda: K.sdf_fwd_grad(pk, pts, st),
      "k3a": lambda: K.sdf_bwd(pk, pts, st, d_sdf, d_grad, None, scratch)}[which]
for _ in range(2):
    fn()
torch.cuda.synchronize()
raw.rnb_trace_set.argtypes = [C.c_void_p]
raw.rnb_trace_set(trace.data_ptr())
fn()
torch.cuda.synchronize()
raw.rnb_trace_set(None)
t = trace.cpu().numpy().reshape(4, 4096, 4)
n_steps = {"k1": 8, "k2": 17, "k3a": 16}[which]
t0 = t[2, 0, 1]
np.save(os.path.join(ROOT, "gpurun_out", f"trace_{which}.npy"), t)
# steady state: tiles 2.. of CTA 0
print(f"{which}: per-step timeline of CTA 0 (cycles), tile 3")
for s in range(3 * n_steps, 4 * n_steps):
    a_wait0, a_seen, acc_commit = t[2, s, :3]
    e_wait0, e_seen, e_sig = t[3, s + 1 if which != "k1" else s, :3] if False else t[3, s, :3]
    print(f"step {s % n_steps:2d}: mma waits a_ready {a_seen - a_wait0:6d} | mma issue phase {acc_commit - a_seen:6d} | "
          f"epi waits acc {e_seen - e_wait0:6d} (acc seen {e_seen - a_seen:6d} after a_ready) | epi work -> signal {e_sig - e_seen:6d}")
# per-slice details of one step
it0 = None
sl = t[1]
print("per-slice (MMA thread): wait_full, then issue; producer: wait_empty, issue  -- relative to a_ready of that step")
step = 3 * n_steps + min(2, n_steps - 1)
a_seen = t[2, step, 1]
# find slices whose full_seen lies between a_seen and acc commit
lo, hi = a_seen, t[2, step, 2]
for i in range(4096):
    if sl[i, 1] >= lo and sl[i, 2] <= hi + 10 and sl[i, 1] > 0:
        p = t[0, i]
        print(f"  slice {i}: producer wait_empty {p[0] - lo:7d}..{p[1] - lo:7d} issued {p[2] - lo:7d} | mma wait_full {sl[i, 0] - lo:6d}..{sl[i, 1] - lo:6d} issued {sl[i, 2] - lo:6d}")
