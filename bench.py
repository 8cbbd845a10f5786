#!/usr/bin/env python
"""bench.py -- throughput of the RNb-NeuS hot path on B200 (contract: see the task prompt / DESIGN.md 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload dp8192|b512|b512_noalbedo|womask_b512|render_bg|grid512|perray]
                    [--impl reference] [--no-extras]

A "step" is one train_rnb iteration over one synthetic ray batch: render_rnb_warmup forward + the loss of
exp_runner.py:241-256 + backward (eikonal double-backward included) + the gradient all-reduce when N > 1 + the Adam
update (exp_runner.py:259-263).  Default workload (BASELINE.json configs[3]): wmask_rnb.conf with the albedo network,
8192 rays per GPU, weak scaling.  The default line also carries the second half of BASELINE.json's metric
("grid512": this rank's x-slab of the 512^3 SDF lattice, configs[4]), the per-ray kernels at a streaming size ("perray"),
the reference on the same GPU through PyTorch-CUDA ("cuda_reference"), the CPU arm ("cpu_baseline") and, under torchrun,
a data-parallel equivalence check ("dp_check").  Prints ONE JSON line on rank 0.
"""
import argparse
import datetime
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rnb-neus-fork_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

# algorithmic FLOPs
FLOP_SDF_ONLY = 918016
FLOP_FWD_FULL = 1049088
FLOP_DX = 917504
FLOP_BWD_DATA = 2 * (458752 + 514560)
FLOP_BWD_DW = 2 * (458752 + 524544)
FLOP_ALB_FWD = 291328
FLOP_ALB_BWD = 873984 - 291328
FLOP_PER_RAY = {True: 967303168, False: 855433216}          # with / without the albedo net

# algorithmic HBM bytes per fine point of the stream-carrying kernels (fp16 streams of 512 B/point,
# 128 B/point for the 64-wide ones; fp32 per-point inputs/outputs).  ncu's dram__bytes for the same launches are in
# profiles/r01_ncu_summary_s4.md (sdf_fwd_grad 11.7 KB/point, sdf_bwd_data 25.6 KB/point measured).
BYTES_PER_POINT = {
    "sdf_fwd_grad": 128 + 8 * 512 + 8 * 512 + 512 + 7 * 512 + 28,           # in0, a_l, w_l, feat written; a_l re-read
    "sdf_bwd_data": (2 * 8 + 8 + 8) * 512 + 512 + 16 + 128 + (8 + 8 + 1) * 512,  # a(x2), w, uin read; uin0, uin, zbar, dfeat written
    "dw_gemm": 1280 + 7 * 4 * 512 + 1024,                                     # four operand streams per layer
    "albedo_fwd": 512 + 128 + 512 + 512 + 36,
    "albedo_bwd": 2 * 512 + 3 * 512 + 40,
}


def source_sha():
    """Hash of the kernel sources this build was made from.  profiles/r02_traffic.json (written by
    profiles/summarize_ncu.py --traffic from an `ncu --set full` capture) carries the hash of the sources it was measured
    on; `roofline.traffic` is printed only when the two agree, so a stale figure can never ride along a changed kernel."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "rnb-neus-fork_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic():
    """-> ({kernel: dram bytes per point}, note)"""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.isfile(p):
        return {}, "no ncu capture committed for this build"
    d = json.load(open(p))
    if d.get("source_sha") != source_sha():
        return {}, f"profiles/r02_traffic.json was captured on sources {d.get('source_sha')}, this build is {source_sha()}"
    return {k: v["dram_bytes_per_point"] for k, v in d["kernels"].items()}, f"profiles/r02_traffic.json ({d.get('capture')})"


WORKLOADS = {
    "dp8192": dict(rays=8192, no_albedo=False, desc="wmask_rnb.conf train_rnb (render_rnb_warmup fwd + loss + bwd + eikonal), "
                   "8192 rays/GPU, 64+64 samples, 3 lights, albedo net on"),
    "b512_noalbedo": dict(rays=512, no_albedo=True, desc="wmask_rnb_noalbedo.conf train_rnb --no_albedo, 512 rays, 64+64 samples"),
    "b512": dict(rays=512, no_albedo=False, desc="wmask_rnb.conf train_rnb, 512 rays, 64+64 samples"),
    "womask_b512": dict(rays=512, no_albedo=False, womask=True,
                        desc="womask_rnb.conf train_rnb as shipped (confs/womask_rnb.conf:28,37,85: n_outside = 0, mask_weight = 0 "
                             "so the mask is all ones, cos_anneal_ratio = iter/50000 < 1, render_rnb after the warm-up), 512 rays. "
                             "render_rnb* with n_outside > 0 has no reference (the reference's own call raises, SURVEY fact 5) "
                             "and is not provided"),
    "render_bg": dict(rays=512, no_albedo=False, render_bg=True,
                      desc="NeuSRenderer.render() with the NeRF++ background field, n_outside = 32 (models/renderer.py:556-648), "
                           "forward only like its one reference caller (render_novel_image), 512 rays x (128 + 32) samples"),
    "grid512": dict(rays=0, no_albedo=True, desc="validate_mesh extract_fields, 512^3 SDF lattice sharded in x-slabs"),
    "perray": dict(rays=131072, no_albedo=False, desc="compositing + hierarchical-sampling kernels alone at 131072 rays "
                   "(HBM-bound per-ray kernels, SURVEY 8d algorithmic bytes)"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def loss_fn(out, true_rgb, mask, igr_weight=0.1, mask_weight=0.1):
    """reference exp_runner.py:241-256 (mask_weight = 0 in womask_rnb.conf: the BCE term drops out)"""
    mask_sum = mask.sum() + 1e-5
    err = ((out["color_fine"] - true_rgb) * mask[None, :, :]).reshape(-1, 3)
    color_loss = F.l1_loss(err, torch.zeros_like(err), reduction="sum") / (mask_sum * true_rgb.shape[0])
    mask_loss = F.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1.0 - 1e-3), mask)
    return color_loss + out["gradient_error"] * igr_weight + mask_loss * mask_weight


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md).  The sampler is started
    BEFORE the warm-up steps (nvidia-smi needs up to a few hundred ms to come up, longer than a short timed region) and
    every row carries nvidia-smi's timestamp; `mark_start()` / `mark_end()` bracket the timed region and only rows inside
    it are used.  If fewer than two rows fall inside (timed region shorter than two sampling periods), the rows of the
    warm-up steps -- same kernels, same load -- are used as well and `window` says so."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.t_begin = datetime.datetime.now()
        self.t0 = self.t1 = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                       "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def mark_start(self):
        self.t0 = datetime.datetime.now()

    def mark_end(self):
        self.t1 = datetime.datetime.now()

    @staticmethod
    def _parse(rows, lo, hi):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                t = datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f")
                if not (lo <= t <= hi):
                    continue
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for i, nme in enumerate(names):
                if len(r) > 4 + i and r[4 + i].strip().lower().startswith("active"):
                    reasons.add(nme)
        return sm, mx, reasons

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        if self.t1 is None:
            self.mark_end()
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        pad = datetime.timedelta(milliseconds=25)
        t0 = self.t0 or self.t_begin
        sm, mx, reasons = self._parse(rows, t0 - pad, self.t1 + pad)
        window = "timed region"
        if len(sm) < 2:
            sm, mx, reasons = self._parse(rows, self.t_begin, self.t1 + pad)
            window = "warm-up + timed region (timed region shorter than two sampling periods)"
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm), window=window)


def build(device, geometric=True):
    from models.fields import NeRF, SDFNetwork, SingleVarianceNetwork, RenderingNetwork
    from models.renderer import NeuSRenderer
    from rnb_b200 import synth
    torch.manual_seed(0)
    conf = synth.WMASK_CONF
    nerf = NeRF(**conf["nerf"])
    sdf = SDFNetwork(**conf["sdf_network"])
    var = SingleVarianceNetwork(**conf["variance_network"])
    col = RenderingNetwork(**conf["rendering_network"])
    for m in (nerf, sdf, var, col):
        m.to(device)
    renderer = NeuSRenderer(nerf, sdf, var, col, **conf["neus_renderer"])
    renderer.color_depth = 3
    return renderer, sdf, var, col


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, wl):
    """The path's CPU implementation on the box's host cores: the unmodified reference when /root/reference is
    present (build container), else oracle/torch_port.py (same algorithm, float32 torch + autograd like the reference).
    A bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_loader
    from rnb_b200 import synth
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    if args.workload == "grid512":
        n_q = 64 ** 3
        sample = f"{n_q} lattice queries (one 64^3 block, the reference's own block size) per step"
    else:
        n_rays = 512
        sample = f"{n_rays} rays of the workload per step (the reference's own batch size; fwd + loss + bwd)"
    kind = "reference" if ref_loader.available() else "port"
    times = []
    if kind == "reference":
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        from oracle.gen_golden import build_reference_nets, loss_fn as ref_loss
        ref, nerf, sdf, var, col = build_reference_nets(False)
        renderer = ref.renderer.NeuSRenderer(nerf, sdf, var, col, **synth.WMASK_CONF["neus_renderer"])
        renderer.color_depth = 3
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            if args.workload == "grid512":
                with torch.no_grad():
                    ref.renderer.extract_fields(torch.tensor([-1.01] * 3), torch.tensor([1.01] * 3), 64, lambda p: -sdf.sdf(p))
            else:
                b = synth.make_batch(n_rays, 3, True, 1, view=it)
                for m in (sdf, var, col):
                    m.zero_grad()
                out = renderer.render_rnb_warmup(b["rays_o"], b["rays_d"], b["near"], b["far"], b["lights_dir"],
                                                 cos_anneal_ratio=1.0, no_albedo=wl["no_albedo"])
                ref_loss(out, b["true_rgb"], b["mask"], 0.1, 0.1, 3)[0].backward()
            if it >= args.warmup:
                times.append(time.perf_counter() - t0)
    else:
        # the reference tree is absent (GPU box): oracle/torch_port.py, the float32 torch + autograd restatement of the same
        # algorithm (pinned against the reference's fixtures in tests/test_oracle_golden.py), on all host threads
        from oracle import torch_port as T
        renderer, sdf, var, col = build("cpu")
        leaf = lambda m: {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
        sdf_sd, col_sd = leaf(sdf), leaf(col)
        variance = var.variance.detach().clone().requires_grad_(True)
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            if args.workload == "grid512":
                T.extract_block(sdf_sd, [-1.01] * 3, [1.01] * 3, 64)
            else:
                b = synth.make_batch(n_rays, 3, True, 1, view=it)
                for t in list(sdf_sd.values()) + list(col_sd.values()) + [variance]:
                    t.grad = None
                out = T.render_rnb(sdf_sd, col_sd, variance, b["rays_o"], b["rays_d"], b["near"], b["far"], b["lights_dir"],
                                   b["t_rand"], 1.0, True, wl["no_albedo"])
                T.loss_fn(out, b["true_rgb"], b["mask"]).backward()
            if it >= args.warmup:
                times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    units = n_q if args.workload == "grid512" else n_rays
    val = units / sec
    unit = "SDF queries/s" if args.workload == "grid512" else "rays/s"
    line = dict(impl="reference", metric=metric_name(args.workload), value=val, unit=unit, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic",
                config=dict(workload=wl["desc"], sample=sample),
                cpu_baseline=dict(value=val, unit=unit, cores=cores, kind=kind, sample=sample),
                e2e=dict(value=val, unit=unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def run_perray(args, wl):
    """The HBM-bound per-ray kernels (K5 compositing fwd/bwd, K6 up-sampling) at a ray count where they stream:
    achieved GB/s = SURVEY 8d algorithmic bytes / cudaEvent time, against the measured copy bandwidth."""
    from rnb_b200 import kernels as K
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    pk = peaks()
    B = wl["rays"]
    g = torch.Generator(device="cpu").manual_seed(1)
    o = torch.nn.functional.normalize(torch.randn(1, 3, generator=g), dim=-1).expand(B, 3).contiguous().to(dev) * 3.0
    d = torch.nn.functional.normalize(-o + 0.3 * torch.randn(B, 3, generator=g).to(dev), dim=-1).contiguous()
    mid = -(o * d).sum(-1, keepdim=True)
    near, far = mid - 1.0, mid + 1.0
    z64 = K.coarse_z(near, far, None, 64)
    z, _ = K.final_merge(torch.sort(torch.cat([z64, z64 + 0.013], -1), -1)[0].contiguous(), None, 2.0 / 64)
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]
    sdf = (pts.norm(dim=-1) - 0.5).reshape(-1).contiguous()
    grad = torch.nn.functional.normalize(pts, dim=-1).reshape(-1, 3).contiguous()
    alb = torch.rand(B * 128, 3, device=dev)
    lights = torch.nn.functional.normalize(torch.randn(3, 1, 1, 3, device=dev), dim=-1)
    var = torch.full((1,), 0.3, device=dev)
    cp = K.composite_params(o, d, z, sdf, grad, alb, lights, var, 1.0, 1, 2.0 / 64)
    d_color = torch.rand(3, B, 3, device=dev) * 1e-4
    d_ws = torch.rand(B, device=dev) * 1e-4
    d_eik, eik_den = torch.full((1,), 0.1, device=dev), torch.full((1,), float(B * 100), device=dev)
    sdf64 = (o[:, None, :] + d[:, None, :] * z64[:, :, None]).norm(dim=-1) - 0.5

    def timed(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    cases = {
        # bytes per ray: SURVEY 8d (fp32): read 128*(sdf 4 + grad 12 + albedo 12 + z 4) + o,d 24; write colour/sums 44 + three [128] arrays
        "composite_fwd": (lambda: K.composite_fwd(cp), 4096 + 24 + 36 + 44 + 3 * 512),
        "composite_bwd": (lambda: K.composite_bwd(cp, d_color, d_ws, d_eik, eik_den, True), 4096 + 24 + 36 + 40 + 128 * 28 + 4),
        "upsample": (lambda: K.upsample_step(o, d, z64, sdf64, 64.0, 16), 2 * 256 + 24 + 64),
        "final_merge": (lambda: K.final_merge(z, None, 2.0 / 64), 512 + 2 * 512),
    }
    kernels, tot_bytes, tot_ms = {}, 0.0, 0.0
    for name, (fn, bpr) in cases.items():
        ms = timed(fn)
        kernels[name] = dict(ms_per_launch=ms, bytes_per_ray=bpr, gbs=bpr * B / ms / 1e6, frac_of_hbm_peak=bpr * B / ms / 1e6 / pk["hbm"])
        if name == "upsample":      # 600 B/ray against ~1.5 k instructions/ray: its roof is the instruction stream, not memory
            kernels[name]["bound"] = "instruction issue (ncu: issue slots 65 % busy, profiles/r02_ncu_summary.md); the HBM fraction is context"
        if name.startswith("composite"):
            tot_bytes += bpr * B
            tot_ms += ms
    val = tot_bytes / tot_ms / 1e6
    line = dict(metric="compositing fwd+bwd GB/s", value=val, unit="GB/s", n_gpus=1, steps=20, warmup=3, ms_per_step=tot_ms,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=wl["desc"], l2="34 KB/ray x 131072 rays = 4.5 GB of inputs+outputs per pass, far above the 126 MB L2"),
                roofline=dict(bound="hbm", kernel="composite_fwd+bwd", achieved=val, peak=pk["hbm"], unit="GB/s", frac=val / pk["hbm"],
                              traffic=None, peak_source=f"MEASURED_PEAKS.json hbm_gbs ({pk['src']})"),
                kernels=kernels, gpu_launches=23 * len(cases))
    print(json.dumps(line))


def metric_name(workload):
    return "mesh SDF queries/s" if workload == "grid512" else "train rays/s (fwd+bwd+eikonal)"


def cpu_baseline(args, wl):
    """Oracle port (or the reference, when present) on the host cores, bounded sample; rank 0, N = 1 only."""
    from oracle import ref_loader
    ns = argparse.Namespace(**vars(args))
    ns.steps, ns.warmup = (2, 1)
    import io
    import contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        run_reference(ns, wl)
    try:
        return json.loads(buf.getvalue().strip().splitlines()[-1])["cpu_baseline"]
    except Exception as e:  # noqa: BLE001
        return dict(value=None, unit="rays/s", cores=os.cpu_count(), kind="port", sample=f"failed: {e}")


# ------------------------------------------------------------------------------------------------ B200 arm
class Ctx:
    """Process-wide handles of one bench run."""

    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.args = args
        self.pk = peaks()

    def sync_all(self):
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            torch.cuda.synchronize()

    def timed(self, fn, steps):
        """K steps between barrier + synchronize on both sides, CUDA events on the launch stream, MAX over ranks."""
        self.sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        self.sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms) / steps


def grid512_section(cx, sdf, steps=3, warmup=1):
    """BASELINE.json configs[4]: this rank's x-slab of the 512^3 lattice (models/renderer.py:10-25, 1219-1224).
    value = whole-lattice queries / slowest rank's time; e2e adds the D2H copy of the slab into pinned host memory."""
    from rnb_b200 import grid, lib as L
    R = 512
    x0, x1 = grid.slab_bounds(R, cx.rank, cx.world)
    out = torch.empty(x1 - x0, R, R, dtype=torch.float32, device=cx.dev)
    host = torch.empty(x1 - x0, R, R, dtype=torch.float32).pin_memory()
    bmin, bmax = [-1.01] * 3, [1.01] * 3
    step = lambda i: grid.sdf_slab(sdf, bmin, bmax, R, x0, x1, out=out)
    for i in range(warmup):
        step(i)
    L.profile_enable(True)
    n0 = L.launch_count()
    ms = cx.timed(step, steps)
    launches = L.launch_count() - n0
    prof = L.profile_collect()
    L.profile_enable(False)

    def e2e_step(i):
        step(i)
        host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    ms_e2e = cx.timed(e2e_step, 2)
    n_launch = (x1 - x0) * R * R
    k_ms = prof["sdf_fwd"][0] / prof["sdf_fwd"][1]
    tf = FLOP_SDF_ONLY * n_launch / k_ms / 1e9
    return dict(metric="mesh SDF queries/s", value=R ** 3 / ms * 1e3, unit="SDF queries/s", ms_per_step=ms, steps=steps,
                warmup=warmup, scaling="strong", gpu_launches=int(launches),
                config=dict(workload=WORKLOADS["grid512"]["desc"], slab=f"x in [{x0},{x1}) of {R}", parallelism=f"slab{cx.world}",
                            l2="no activation streams: the kernel reads 2 MB of packed weights (L2-resident) and writes 4 B per "
                               "query; the slab output (>= 64 MiB) is larger than L2"),
                e2e=dict(value=R ** 3 / ms_e2e * 1e3, unit="SDF queries/s", h2d_bytes_per_step=0, d2h_bytes_per_step=n_launch * 4,
                         ms_per_step=ms_e2e),
                roofline=dict(bound="tensor", kernel="sdf_fwd", achieved=tf, peak=cx.pk["tf_sustained"], unit="TFLOP/s",
                              frac=tf / cx.pk["tf_sustained"], frac_of_burst_peak=tf / cx.pk["tf_burst"],
                              algorithmic_flop_per_query=FLOP_SDF_ONLY, queries_per_launch=n_launch, ms_per_launch=k_ms,
                              hbm_frac=4.0 * n_launch / k_ms / 1e6 / cx.pk["hbm"], traffic=None,
                              peak_source=f"MEASURED_PEAKS.json bf16_tflops_sustained ({cx.pk['src']})"))


def perray_section(cx, B=131072):
    """K5 / K6 at a ray count where they stream (the 8192-ray launches of the train step are single waves)."""
    import io
    import contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        run_perray(cx.args, dict(WORKLOADS["perray"], rays=B))
    d = json.loads(buf.getvalue().strip().splitlines()[-1])
    return dict(rays=B, kernels=d["kernels"], roofline=d["roofline"], config=d["config"])


def cuda_reference_section(cx, sizes=(512, 2048), steps=3):
    """The reference's own implementation through PyTorch-CUDA (ATen + cuBLAS + autograd, fp32, TF32 off) on the same GPU:
    the unmodified reference when its tree is staged (oracle/stage_reference.py -> baseline/_ref, or /root/reference),
    else oracle/torch_port.py.  A GPU-vs-GPU baseline next to the CPU one ("optional second baseline")."""
    from oracle import ref_loader
    from rnb_b200 import synth
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    kind = "reference" if ref_loader.available() else "port"
    res = dict(kind=kind, dtype="f32 (allow_tf32 = False)", steps=steps, sizes={})
    try:
        if kind == "reference":
            from oracle.gen_golden import build_reference_nets, loss_fn as ref_loss
            ref, nerf, sdf, var, col = build_reference_nets(False)
            for m in (nerf, sdf, var, col):
                m.to(cx.dev)
            renderer = ref.renderer.NeuSRenderer(nerf, sdf, var, col, **synth.WMASK_CONF["neus_renderer"])
            renderer.color_depth = 3
        else:
            from oracle import torch_port as T
            _, sdf, var, col = build("cpu")
            leaf = lambda m: {k: v.detach().clone().to(cx.dev).requires_grad_(True) for k, v in m.state_dict().items()}
            sdf_sd, col_sd = leaf(sdf), leaf(col)
            variance = var.variance.detach().clone().to(cx.dev).requires_grad_(True)
        batches = {B: {k: v.to(cx.dev) for k, v in synth.make_batch(B, 3, True, 1).items()} for B in sizes}
        torch.set_default_tensor_type("torch.cuda.FloatTensor")
        for B in sizes:
            b = batches[B]

            def step(i):
                # the reference builds its temporaries with bare factory calls and relies on exp_runner.py:669's global
                # torch.set_default_tensor_type('torch.cuda.FloatTensor') (which the legacy torch.Tensor([...]) honours too)
                with torch.device(cx.dev):
                    if kind == "reference":
                        for m in (sdf, var, col):
                            m.zero_grad()
                        out = renderer.render_rnb_warmup(b["rays_o"], b["rays_d"], b["near"], b["far"], b["lights_dir"],
                                                         cos_anneal_ratio=1.0, no_albedo=False)
                        ref_loss(out, b["true_rgb"], b["mask"], 0.1, 0.1, 3)[0].backward()
                    else:
                        for t in list(sdf_sd.values()) + list(col_sd.values()) + [variance]:
                            t.grad = None
                        out = T.render_rnb(sdf_sd, col_sd, variance, b["rays_o"], b["rays_d"], b["near"], b["far"],
                                           b["lights_dir"], b["t_rand"], 1.0, True, False)
                        T.loss_fn(out, b["true_rgb"], b["mask"]).backward()
            step(0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                step(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            res["sizes"][str(B)] = dict(ms_per_step=ms, rays_per_s=B / ms * 1e3,
                                        peak_mem_gb=torch.cuda.max_memory_allocated(cx.dev) / 2 ** 30)
        res["value"] = max(v["rays_per_s"] for v in res["sizes"].values())
        res["unit"] = "rays/s"
    except Exception as e:  # noqa: BLE001  (a baseline that cannot run must not take the bench line down with it)
        res["error"] = f"{type(e).__name__}: {e}"[:300]
    finally:
        torch.set_default_tensor_type("torch.FloatTensor")
    torch.cuda.empty_cache()
    return res


def dp_check(cx, train_step, red, batches, exact=None):
    """Data-parallel equivalence with the real kernels under NCCL (reference seam exp_runner.py:170, 194):
    one untimed step -- every rank keeps its pre-reduction flat gradient, all ranks gather them, and the all-reduced
    buffer must equal their mean in the rank order NCCL is NOT obliged to use, so the comparison is to fp32 rounding of a
    sum of `world` terms (bit-exact for world = 2); the ranks must also have drawn different rays."""
    dist = cx.dist
    b = batches[0]
    train_step(b, reduce=False)
    local = red.collect().clone()
    gathered = [torch.empty_like(local) for _ in range(cx.world)]
    dist.all_gather(gathered, local)
    red.all_reduce()
    mean = torch.stack(gathered).double().mean(0)
    got = red.flat.double()
    err = float((got - mean).norm() / mean.norm().clamp_min(1e-300))
    rays = [torch.empty_like(b["rays_d"]) for _ in range(cx.world)]
    dist.all_gather(rays, b["rays_d"].contiguous())
    distinct = all(not torch.equal(rays[0], r) for r in rays[1:])
    differ = float((gathered[0] - gathered[-1]).norm() / gathered[0].norm().clamp_min(1e-30))
    ok = err < 1e-6 and distinct and differ > 1e-3
    res = dict(status="ok" if ok else "FAILED", allreduce_vs_mean_of_ranks_rel_l2=err, ranks_drew_distinct_rays=distinct,
               rank_gradients_differ_rel_l2=differ, flat_buffer_bytes=red.nbytes)
    if exact is not None:
        res["exact_global_batch"] = exact()
        if res["exact_global_batch"]["rel_l2_vs_single_process_batch"] > 5e-3:
            res["status"] = "FAILED"
    return res


def exact_batch_check(cx, renderer, red, no_albedo, rays=256):
    """parallel.ExactBatch under NCCL with the real kernels: `world` ranks x `rays` rays (eikonal numerator / count and
    mask_sum summed over the ranks, gradient shares summed) against ONE process rendering the concatenated batch with the
    reference's loss.  Jitter off (perturb_overwrite = 0) so both see the same sample depths."""
    from rnb_b200 import synth
    from rnb_b200.parallel import ExactBatch
    dist = cx.dist
    b = {k: v.to(cx.dev) for k, v in synth.make_batch(rays, 3, True, 100 + cx.rank).items()}
    keys = ("rays_o", "rays_d", "near", "far", "true_rgb", "mask")
    big = {}
    for k in keys:
        parts = [torch.empty_like(b[k]) for _ in range(cx.world)]
        dist.all_gather(parts, b[k].contiguous())
        big[k] = torch.cat(parts, 1 if k == "true_rgb" else 0)
    lights = b["lights_dir"].clone()
    dist.broadcast(lights, 0)                      # warm-up mode: one light set per view, shared by all rays
    # (1) one process, the whole batch, the reference's loss
    red.zero()
    out = renderer.render_rnb_warmup(big["rays_o"], big["rays_d"], big["near"], big["far"], lights, perturb_overwrite=0,
                                     cos_anneal_ratio=1.0, no_albedo=no_albedo)
    loss_big = loss_fn(out, big["true_rgb"], big["mask"])
    loss_big.backward()
    g_big = red.collect().clone()
    # (2) ExactBatch: every rank its own rays
    eb = ExactBatch(renderer)
    red.zero()
    out = renderer.render_rnb_warmup(b["rays_o"], b["rays_d"], b["near"], b["far"], lights, perturb_overwrite=0,
                                     cos_anneal_ratio=1.0, no_albedo=no_albedo)
    share = eb.loss(out, b["true_rgb"], b["mask"])
    share.backward()
    g_dp = red.all_reduce_sum().clone()
    total = eb.total(share, out)
    renderer.dp_exact_group = None
    rel = float((g_dp - g_big).norm() / g_big.norm().clamp_min(1e-30))
    return dict(rays_per_rank=rays, rel_l2_vs_single_process_batch=rel, loss_single_process=float(loss_big),
                loss_from_shares=float(total))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dp8192", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="train line only: skip grid512 / perray / cuda_reference")
    ap.add_argument("--graph", action="store_true", help="replay the step as one CUDA graph (rnb_b200.graph_step)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    args.warmup = max(args.warmup, 3)
    if args.workload == "perray":
        if int(os.environ.get("RANK", "0")) == 0:
            run_perray(args, wl)
        return

    cx = Ctx(args)
    dist, world, rank, dev, pk = cx.dist, cx.world, cx.rank, cx.dev, cx.pk
    from rnb_b200 import synth, lib as L
    from rnb_b200.parallel import FlatGradAllReducer
    from rnb_b200.optim import FlatAdam

    renderer, sdf, var, col = build(dev)

    if args.workload == "grid512":
        clk = ClockSampler(cx.local)
        clk.mark_start()
        sec = grid512_section(cx, sdf, steps=max(3, args.steps), warmup=args.warmup)
        clk.mark_end()
        line = dict(sec, n_gpus=world, higher_is_better=True, vs_baseline=None, dtype="f16 operands / f32 accumulate",
                    data="synthetic", clocks=clk.stop())
        if rank == 0:
            if world == 1 and not args.no_cpu_baseline:
                line["cpu_baseline"] = cpu_baseline(args, wl)
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return

    B = wl["rays"]
    no_albedo = wl["no_albedo"]
    womask = bool(wl.get("womask"))
    render_bg = bool(wl.get("render_bg"))
    warm = not womask
    if render_bg:
        renderer.n_outside = 32
    params = list(sdf.parameters()) + list(var.parameters()) + ([] if no_albedo else list(col.parameters()))
    red = FlatGradAllReducer(params)
    opt = FlatAdam(params, lr=5e-4, reducer=red)   # one launch over the flat buffers the all-reduce uses
    host_b = [{k: v.pin_memory() for k, v in synth.make_batch(B, 3, warm, 1 + rank, view=i).items()} for i in range(4)]
    if womask:       # mask_weight = 0 -> mask = ones (exp_runner.py:187-194)
        for hb in host_b:
            hb["mask"].fill_(1.0)
    dev_b = [{k: v.to(dev) for k, v in hb.items()} for hb in host_b]
    keys = ("rays_o", "rays_d", "near", "far", "lights_dir", "true_rgb", "mask")
    anneal = 0.3 if womask else 1.0
    mask_w = 0.0 if womask else 0.1

    def fwd_bwd(b, reduce=True):
        # every step starts from weights an optimiser has just updated: bump the version counters so that nothing
        # keyed on them (the cached packed weights of the no_grad sampling pass) carries over from the previous step
        torch.autograd.graph.increment_version(params)
        red.zero()
        fn = renderer.render_rnb_warmup if warm else renderer.render_rnb
        out = fn(b["rays_o"], b["rays_d"], b["near"], b["far"], b["lights_dir"], cos_anneal_ratio=anneal, no_albedo=no_albedo)
        loss = loss_fn(out, b["true_rgb"], b["mask"], mask_weight=mask_w)
        loss.backward()
        if reduce:
            red.all_reduce()
        return loss

    if render_bg:
        def fwd_bwd(b, reduce=True):      # noqa: F811  forward only: render() has no training caller upstream
            with torch.no_grad():
                out = renderer.render(b["rays_o"], b["rays_d"], b["near"], b["far"], cos_anneal_ratio=1.0)
            return out["color_fine"].sum()
    if args.graph:
        from rnb_b200.graph_step import GraphedTrainStep
        gs = GraphedTrainStep(renderer, params, lambda o, rgb, m: loss_fn(o, rgb, m, mask_weight=mask_w), dev_b[0], warmup=warm,
                              no_albedo=no_albedo, reducer=red)

        def fwd_bwd(b, reduce=True):       # noqa: F811  (same contract: grads in the flat buffer, all-reduce outside the graph)
            loss = gs(b)
            if reduce:
                red.all_reduce()
            return loss

    def train(b):
        """one train_rnb iteration: forward + loss + backward (+ all-reduce) + Adam (exp_runner.py:222-263)"""
        loss = fwd_bwd(b)
        if not render_bg:
            opt.step()
        return loss

    step = lambda i: train(dev_b[i % 4])
    clk = ClockSampler(cx.local)
    for i in range(args.warmup):
        step(i)
    check = dp_check(cx, fwd_bwd, red, dev_b, exact=(lambda: exact_batch_check(cx, renderer, red, no_albedo))
                     if not args.graph else None) if (world > 1 and not render_bg) else None
    L.profile_enable(True)
    n0 = L.launch_count()
    cx.sync_all()
    clk.mark_start()
    ms = cx.timed(step, args.steps)
    clk.mark_end()
    launches = L.launch_count() - n0
    if args.graph:      # replays issue no calls through the C-ABI: count the library kernels captured in the graph
        launches += gs.launches_per_replay * args.steps
    prof = L.profile_collect()
    L.profile_enable(False)
    clocks = clk.stop()

    def e2e_step(i):
        hb = host_b[i % 4]
        b = {k: hb[k].to(dev, non_blocking=True) for k in keys}
        loss = train(b)
        return float(loss.detach())              # device -> host read of the step's result
    ms_e2e = cx.timed(e2e_step, args.steps)
    ms_fwd_bwd = cx.timed(lambda i: fwd_bwd(dev_b[i % 4]), max(3, args.steps // 2))
    units = B * world
    h2d = sum(host_b[0][k].numel() * 4 for k in keys)
    value = units / ms * 1e3

    # ---- per-kernel device times (cudaEvent brackets recorded by the library inside the timed region) and rooflines
    nf, nc = B * 128, B * 112
    flops = {"sdf_fwd_grad": (FLOP_FWD_FULL + FLOP_DX) * nf, "sdf_bwd_data": FLOP_BWD_DATA * nf,
             "sdf_bwd_fused": (FLOP_BWD_DATA + FLOP_BWD_DW) * nf, "albedo_fwd": FLOP_ALB_FWD * nf, "albedo_bwd": 2 * 145664 * nf}
    # dw_gemm launches: one for the SDF net, one for the albedo net per step
    flops["dw_gemm"] = (FLOP_BWD_DW * nf + (0 if no_albedo else 2 * 145664 * nf)) / (1 if no_albedo else 2)
    traffic_pp, traffic_note = ncu_traffic()
    kernels = {}
    for name, (tot, cnt) in prof.items():
        d = dict(ms_per_launch=tot / cnt, launches_per_step=cnt / args.steps, share_of_step=tot / (ms * args.steps))
        if name in flops or name == "sdf_fwd":
            fl = FLOP_SDF_ONLY * nc / (cnt / args.steps) if name == "sdf_fwd" else flops[name]
            d["tflops"] = fl / (tot / cnt) / 1e9
            d["frac_of_tensor_peak"] = d["tflops"] / pk["tf_sustained"]
        if name in BYTES_PER_POINT:
            nbytes = BYTES_PER_POINT[name] * nf
            if name == "dw_gemm":       # two launches per step (SDF net, albedo net): the mean launch
                nbytes = (BYTES_PER_POINT[name] + (0 if no_albedo else 2816)) * nf / (1 if no_albedo else 2)
            d["algorithmic_bytes"] = nbytes
            d["gbs"] = nbytes / (tot / cnt) / 1e6
            d["frac_of_hbm_peak"] = d["gbs"] / pk["hbm"]
        if name in traffic_pp:
            d["ncu_dram_bytes"] = traffic_pp[name] * nf
        kernels[name] = d
    cand = [k for k in kernels if "tflops" in kernels[k]]
    top = max(cand, key=lambda k: kernels[k]["share_of_step"]) if cand else None
    roofline = None
    if top:
        k = kernels[top]
        # the MLP kernels are dense contractions, their algorithmic unit is FLOPs and the roof the (sustained)
        # tensor peak.  What keeps this kernel under that roof is its HBM stream traffic: reported beside it.
        roofline = dict(bound="tensor", kernel=top, achieved=k["tflops"], peak=pk["tf_sustained"], unit="TFLOP/s",
                        frac=k["frac_of_tensor_peak"], frac_of_burst_peak=k["tflops"] / pk["tf_burst"],
                        ms_per_launch=k["ms_per_launch"], points_per_launch=nf,
                        algorithmic_flop_per_point=flops[top] / nf if top in flops else FLOP_SDF_ONLY,
                        algorithmic_bytes=k.get("algorithmic_bytes"), hbm_frac=k.get("frac_of_hbm_peak"),
                        traffic=k.get("ncu_dram_bytes"), traffic_source=traffic_note,
                        peak_source=f"MEASURED_PEAKS.json bf16_tflops_sustained ({pk['src']})")
    step_tf = (FLOP_PER_RAY[not no_albedo] * B / ms / 1e9) if not render_bg else None
    line = dict(metric=metric_name(args.workload), value=value, unit="rays/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f16 operands / f32 accumulate",
                data="synthetic",
                config=dict(workload=wl["desc"] + (" [one CUDA graph per step]" if args.graph else ""),
                            step="forward + loss + backward + all-reduce + Adam (FlatAdam, one launch)" if not render_bg else "forward",
                            weights="geometric init, torch.manual_seed(0)",
                            l2="per-step working set (fp16 activation streams, > 1 GB at 8192 rays) exceeds the 126 MB L2; "
                               "4 input batches rotate", parallelism=f"dp{world}"),
                clocks=clocks, e2e=dict(value=units / ms_e2e * 1e3, unit="rays/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=4,
                                        ms_per_step=ms_e2e),
                gpu_launches=int(launches), roofline=roofline, kernels=kernels, fwd_bwd_ms=ms_fwd_bwd, source_sha=source_sha())
    if step_tf is not None:
        line["roofline_step"] = dict(bound="tensor", achieved=step_tf, peak=pk["tf_sustained"], unit="TFLOP/s",
                                     frac=step_tf / pk["tf_sustained"], algorithmic_flop_per_ray=FLOP_PER_RAY[not no_albedo],
                                     floor_ms_at_peak=FLOP_PER_RAY[not no_albedo] * B / pk["tf_sustained"] / 1e9)
    if check is not None:
        line["dp_check"] = check
    extras = args.workload == "dp8192" and not args.no_extras and not args.graph
    if extras:
        line["grid512"] = grid512_section(cx, sdf)
        if rank == 0:
            line["perray"] = perray_section(cx)
        cx.sync_all()
    if rank == 0:
        if world == 1 and extras:
            line["cuda_reference"] = cuda_reference_section(cx)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, wl)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
