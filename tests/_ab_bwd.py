"""A/B of the SDF backward on one box: two-kernel path (RNB_BWD_FUSED=0) vs the fused launch, same inputs.
usage: python tests/_ab_bwd.py [rays] [repl ...]   (each repl = RNB_DW_REPL string to try, e.g. 3,4,4,4,4,4,4,4,2)"""
import os
import sys
sys.path[:0] = ['.', 'tests', 'rnb-neus-fork_b200']
import torch
from test_gpu_e2e import make_renderer
from rnb_b200 import synth, kernels as K, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
repls = sys.argv[2:] or ["3,4,4,4,4,4,4,4,2"]
renderer, sdf, var, col = make_renderer(True)


ONCE = bool(os.environ.get("RNB_AB_ONCE"))       # one launch per configuration (for ncu)


def timeit(fn, n=6, warm=2):
    if ONCE:
        return float("nan")
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


b = {k: v.cuda() for k, v in synth.make_batch(B, 3, True, 1).items()}
pk = ops.packed_sdf_nograd(sdf)
z, mid = ops.hierarchical_sample(sdf, b["rays_o"], b["rays_d"], b["near"], b["far"], b["t_rand"], 64, 64, 4)
pts = K.points_rays(b["rays_o"], b["rays_d"], mid)
st = K.SdfStreams(pts.n_pts, "cuda")
K.sdf_fwd_grad(pk, pts, st)
n = pts.n_pts
g = torch.Generator(device="cuda").manual_seed(3)
d_sdf = torch.randn(n, device="cuda", generator=g) * 1e-4
d_grad = torch.randn(n, 3, device="cuda", generator=g) * 1e-5
d_feat = torch.randn(n, 256, device="cuda", generator=g) * 1e-7
scratch = torch.empty(K.L.load().rnb_sdf_bwd_scratch_bytes(n), dtype=torch.uint8, device="cuda")


def run():
    return K.sdf_bwd(pk, pts, st, d_sdf, d_grad, d_feat, scratch)


os.environ["RNB_BWD_FUSED"] = "0"
ref = run()
torch.cuda.synchronize()
ref = ([t.clone() for t in ref[0]], [t.clone() for t in ref[1]])
ms0 = timeit(run)
print(f"B={B} points={n}: two-kernel path {ms0:.3f} ms")
os.environ["RNB_BWD_FUSED"] = "1"
for r in repls:
    if ":" in r:
        r, stg = r.split(":")
        os.environ["RNB_FUSED_STAGGER_US"] = stg
    os.environ["RNB_DW_REPL"] = r
    out = run()
    torch.cuda.synchronize()
    worst = 0.0
    for l in range(9):
        for a, c in ((out[0][l], ref[0][l]), (out[1][l], ref[1][l])):
            e = float((a - c).norm() / c.norm().clamp_min(1e-30))
            worst = max(worst, e)
    ms1 = timeit(run)
    os.environ["RNB_FUSED_DBG"] = "1"
    run()
    torch.cuda.synchronize()
    del os.environ["RNB_FUSED_DBG"]
    off = K.L.load().rnb_sdf_bwd_debug_offset(n)
    dbg = scratch[off:off + 2 * 160 * 8].view(torch.int64).view(2, 160).cpu()
    nw = sum(int(x) for x in r.split(','))
    t0 = int(dbg[1, :148].min())
    end = (dbg[0, :148] - t0).double() / 1e6
    layers = [l for l, c in enumerate(int(x) for x in r.split(',')) for _ in range(c)]
    per_layer = {}
    for j, l in enumerate(layers):
        per_layer.setdefault(l, []).append(round(float(end[j]), 2))
    print(f"      chain blocks end at {float(end[nw:].min()):.2f}..{float(end[nw:].max()):.2f} ms; workers per layer: {per_layer}")
    print(f"   stagger {os.environ.get('RNB_FUSED_STAGGER_US')} fused repl={r} (workers {sum(int(x) for x in r.split(','))}): {ms1:.3f} ms   worst rel-L2 vs two-kernel {worst:.2e}")
