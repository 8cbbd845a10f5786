"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/rnb_b200.h declares, the drop-in
modules keep the reference's constructor / state_dict contract, the product path refuses to run without CUDA, and
the multi-process host logic (gradient all-reduce, slab partition) works over gloo with world_size 2."""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "rnb_b200.h")).read()
    return sorted(set(re.findall(r"RNB_API\s+[\w\s\*]+?\b(rnb_\w+)\s*\(", src)))


def test_cabi_exports_every_declared_symbol():
    from rnb_b200 import lib
    so = lib.LIB_PATH
    if not os.path.isfile(so):
        import __graft_entry__ as ge
        ge.build()
    handle = lib.load()
    declared = header_symbols()
    assert len(declared) >= 25
    nm = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (rnb_\w+)", nm))
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    for s in declared:
        assert hasattr(handle, s)
    # every symbol the Python binding uses is declared in the header
    assert set(lib.exported_symbols()) <= set(declared)
    assert handle.rnb_version() >= 100
    assert handle.rnb_padded_points(1) == 128 and handle.rnb_padded_points(129) == 256
    assert handle.rnb_stream_bytes(1000, 256) == 1024 * 512
    assert handle.rnb_sdf_wblob_bytes() == 2162688


def test_no_tensor_core_fallbacks_in_binary():
    """The library is sm_100a only and its MLP kernels are tcgen05 / TMA-engine code (SASS mnemonics)."""
    from rnb_b200 import lib
    try:
        sass = subprocess.run(["cuobjdump", "-sass", lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    except (OSError, subprocess.CalledProcessError):
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass and "LDTM" in sass and "UBLKCP" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", "")          # no legacy mma.sync path
    assert "FFMA2" in sass and "FMUL2" in sass                   # packed-fp32 epilogues (Blackwell two-lane fp32)


def test_modules_match_reference_contract():
    from models import fields
    from rnb_b200 import synth
    conf = synth.WMASK_CONF
    torch.manual_seed(0)
    sdf = fields.SDFNetwork(**conf["sdf_network"])
    col = fields.RenderingNetwork(**conf["rendering_network"])
    nerf = fields.NeRF(**conf["nerf"])
    var = fields.SingleVarianceNetwork(**conf["variance_network"])
    keys = list(sdf.state_dict())
    assert keys[:3] == ["lin0.bias", "lin0.weight_g", "lin0.weight_v"]
    shapes = {k: tuple(v.shape) for k, v in sdf.state_dict().items()}
    assert shapes["lin0.weight_v"] == (256, 39) and shapes["lin3.weight_v"] == (217, 256) and shapes["lin8.weight_v"] == (257, 256)
    assert shapes["lin8.weight_g"] == (257, 1)
    assert sum(p.numel() for p in sdf.parameters()) == 529076
    assert sum(p.numel() for p in col.parameters()) == 146694
    assert sum(p.numel() for p in nerf.parameters()) == 606596
    assert tuple(col.state_dict()["lin0.weight_v"].shape) == (256, 310)
    assert list(var.state_dict()) == ["variance"] and abs(float(var.variance) - 0.3) < 1e-7
    assert "pts_linears.5.weight" in nerf.state_dict() and tuple(nerf.state_dict()["pts_linears.5.weight"].shape) == (256, 340)
    # geometric init: sdf ~ |x| - r at init; last layer mean sqrt(pi)/sqrt(256), bias -0.5
    assert abs(float(sdf.lin8.bias[0]) + 0.5) < 1e-6
    eff = sdf.effective_weights()
    assert len(eff) == 9 and tuple(eff[4][0].shape) == (256, 256)
    # weight-norm folding is exactly g * v / |v|
    W0 = eff[0][0]
    v, g = sdf.lin0.weight_v, sdf.lin0.weight_g
    assert torch.allclose(W0, g * v / v.norm(dim=1, keepdim=True), atol=1e-7)


def test_reference_checkpoints_load_both_ways():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present on this machine")
    from models import fields
    from rnb_b200 import synth
    ref = ref_loader.load()
    conf = synth.WMASK_CONF
    torch.manual_seed(3)
    theirs = ref.fields.SDFNetwork(**conf["sdf_network"])
    torch.manual_seed(4)
    mine = fields.SDFNetwork(**conf["sdf_network"])
    mine.load_state_dict(theirs.state_dict())
    theirs2 = ref.fields.SDFNetwork(**conf["sdf_network"])
    theirs2.load_state_dict(mine.state_dict())
    for (k1, v1), (k2, v2) in zip(theirs.state_dict().items(), theirs2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    for cls, key in (("RenderingNetwork", "rendering_network"), ("NeRF", "nerf")):
        a = getattr(ref.fields, cls)(**conf[key])
        b = getattr(fields, cls)(**conf[key])
        b.load_state_dict(a.state_dict())


def test_product_path_fails_loudly_without_cuda():
    from models import fields
    from models.renderer import NeuSRenderer
    from rnb_b200 import synth
    conf = synth.WMASK_CONF
    sdf = fields.SDFNetwork(**conf["sdf_network"])
    with pytest.raises(RuntimeError, match="CUDA"):
        sdf.sdf(torch.zeros(4, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        sdf(torch.zeros(4, 3))
    r = NeuSRenderer(fields.NeRF(**conf["nerf"]), sdf, fields.SingleVarianceNetwork(0.3),
                     fields.RenderingNetwork(**conf["rendering_network"]), **conf["neus_renderer"])
    b = synth.make_batch(4)
    with pytest.raises(RuntimeError, match="CUDA"):
        r.render_rnb_warmup(b["rays_o"], b["rays_d"], b["near"], b["far"], b["lights_dir"])
    # nothing under the product tree imports the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rnb-neus-fork_b200")):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("the oracle", ""), os.path.join(dirpath, f)


def test_synthetic_batch_shapes_and_determinism():
    from rnb_b200 import synth
    a, b = synth.make_batch(64, 3, True, 1), synth.make_batch(64, 3, True, 1)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert a["lights_dir"].shape == (3, 1, 1, 3) and synth.make_batch(8, 3, False)["lights_dir"].shape == (3, 8, 1, 3)
    assert torch.allclose(a["rays_d"].norm(dim=-1), torch.ones(64), atol=1e-6)
    assert torch.allclose(a["rays_o"].norm(dim=-1), torch.full((64,), 3.0), atol=1e-5)
    assert torch.allclose(a["far"] - a["near"], torch.full((64, 1), 2.0), atol=1e-6)


def test_slab_partition_covers_lattice():
    from rnb_b200 import grid
    for R, world in ((512, 8), (512, 3), (128, 1), (33, 4)):
        spans = [grid.slab_bounds(R, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == R
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0 and a0 <= a1


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "rnb-neus-fork_b200"))
    from rnb_b200.parallel import FlatGradAllReducer, rank_seed
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Softplus(beta=100), torch.nn.Linear(7, 1))
    red = FlatGradAllReducer(list(net.parameters()))
    g = torch.Generator().manual_seed(rank_seed(3, rank, world))
    x = torch.randn(16, 5, generator=g)
    for _ in range(2):                      # second round checks zero() + in-place accumulation
        red.zero()
        net(x).pow(2).mean().backward()
        red.all_reduce()
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(red.params, red.views))
    assert all(v.data_ptr() % 16 == 0 for v in red.views)          # slices are padded to 16-byte boundaries
    q.put((rank, torch.cat([v.reshape(-1) for v in red.views]).numpy(), x.numpy()))
    dist.destroy_process_group()


def test_data_parallel_allreduce_gloo_world2():
    """R ranks x B rays reproduce the single-process gradient of the mean of the per-rank losses (DDP semantics)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert np.allclose(res[0][1], res[1][1])                   # identical on both ranks
    assert not np.allclose(res[0][2], res[1][2])               # different rays per rank (rank_seed)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Softplus(beta=100), torch.nn.Linear(7, 1))
    loss = sum(net(torch.from_numpy(r[2])).pow(2).mean() for r in res) / 2
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).numpy()
    assert np.allclose(res[0][1], ref, atol=1e-6)


def _exact_worker(rank, world, port, q):
    """ExactBatch host logic on CPU: a stand-in 'renderer' (tiny differentiable net) whose eikonal-like term is a ratio of
    sums over the batch; the per-rank loss shares, gradient shares summed with all_reduce_sum."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "rnb-neus-fork_b200"))
    from rnb_b200.parallel import ExactBatch, FlatGradAllReducer
    torch.manual_seed(0)
    net = torch.nn.Linear(4, 5)
    red = FlatGradAllReducer(list(net.parameters()))
    g = torch.Generator().manual_seed(10 + rank)
    B = 6 + rank                                         # unequal batches: the normalisers really are global
    x = torch.randn(B, 4, generator=g)
    true_rgb = torch.rand(3, B, 3, generator=g)
    mask = (torch.rand(B, 1, generator=g) < 0.6).float()

    class R:
        dp_exact_group = None
    eb = ExactBatch(R)
    h = net(x)
    num, den = (h[:, 4] ** 2).sum(), torch.tensor(float(B))
    both = torch.stack([num.detach(), den])
    dist.all_reduce(both)
    eik = num / (both[1] + 1e-5) + (both[0] - num.detach()) / (both[1] + 1e-5)      # global ratio, differentiable through local points
    out = dict(color_fine=h[:, :3].sigmoid()[None].expand(3, B, 3), weight_sum=h[:, 3:4].sigmoid(), gradient_error=eik)
    red.zero()
    share = eb.loss(out, true_rgb, mask)
    share.backward()
    flat = red.all_reduce_sum().clone()
    total = eb.total(share, out)
    q.put((rank, flat.numpy(), x.numpy(), true_rgb.numpy(), mask.numpy(), float(total)))
    dist.destroy_process_group()


def test_exact_batch_shares_sum_to_the_global_batch_gradient_gloo_world2():
    """parallel.ExactBatch: per-rank loss shares with global normalisers (mask_sum, ray count, eikonal denominator) + SUM of
    the gradient shares == the reference loss (exp_runner.py:241-256) on the concatenated batch, also for unequal batches."""
    import torch.multiprocessing as mp
    import torch.nn.functional as F
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_exact_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert np.allclose(res[0][1], res[1][1])
    torch.manual_seed(0)
    net = torch.nn.Linear(4, 5)
    x = torch.from_numpy(np.concatenate([r[2] for r in res], 0))
    rgb = torch.from_numpy(np.concatenate([r[3] for r in res], 1))
    mask = torch.from_numpy(np.concatenate([r[4] for r in res], 0))
    h = net(x)
    B = x.shape[0]
    color = h[:, :3].sigmoid()[None].expand(3, B, 3)
    err = ((color - rgb) * mask[None]).reshape(-1, 3)
    loss = (F.l1_loss(err, torch.zeros_like(err), reduction="sum") / ((mask.sum() + 1e-5) * 3)
            + 0.1 * (h[:, 4] ** 2).sum() / (B + 1e-5) + 0.1 * F.binary_cross_entropy(h[:, 3:4].sigmoid().clip(1e-3, 1 - 1e-3), mask))
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).numpy()
    got = res[0][1][:ref.size]
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-7), np.abs(got - ref).max()
    assert abs(res[0][5] - float(loss)) < 1e-5 * abs(float(loss))


def _cpu_triangle_soup(u, x_lo, x_hi):
    """marching cubes with the generated tables on the cells x_lo <= x < x_hi of a small host lattice (test helper)"""
    from rnb_b200 import mc_tables as M
    nx, ny, nz = u.shape
    verts, keys = [], []
    for x in range(x_lo, x_hi):
        for y in range(ny - 1):
            for z in range(nz - 1):
                case = sum(1 << i for i, (ox, oy, oz) in enumerate(M.CORNER_OFFSETS) if u[x + ox, y + oy, z + oz] > 0)
                for e in M.TRI_TABLE[case][:3 * M.TRI_COUNT[case]]:
                    a, b = M.EDGE_CORNERS[e]
                    pa, pb = np.array([x, y, z]) + M.CORNER_OFFSETS[a], np.array([x, y, z]) + M.CORNER_OFFSETS[b]
                    t = (0.0 - u[tuple(pa)]) / (u[tuple(pb)] - u[tuple(pa)])
                    verts.append(pa + t * (pb - pa))
                    keys.append(((pa[0] * ny + pa[1]) * nz + pa[2]) * 3 + int(e) // 4)
    return torch.tensor(np.array(verts).reshape(-1, 3), dtype=torch.float32), torch.tensor(keys, dtype=torch.int64)


def _mesh_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "rnb-neus-fork_b200"))
    from rnb_b200 import grid
    n = 12
    g = np.linspace(-1, 1, n)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    u = 0.6 - np.sqrt(X ** 2 + Y ** 2 + Z ** 2)
    x0, x1 = grid.slab_bounds(n, rank, world)
    verts, keys = _cpu_triangle_soup(u, x0, min(x1, n - 1))       # cells of this rank's slab (the overlap plane closes the seam)
    out = grid.gather_and_weld(verts, keys)
    q.put((rank, None if out is None else (out[0], out[1])))
    dist.destroy_process_group()


def test_sharded_mesh_gather_and_weld_gloo_world2():
    """slab-wise marching cubes + one gather of the triangles == one pass over the whole lattice (closed, genus 0)"""
    import torch.multiprocessing as mp
    from collections import Counter
    sys.path.insert(0, os.path.join(ROOT, "rnb-neus-fork_b200"))
    from rnb_b200 import grid
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_mesh_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[1] is None
    V, T = res[0]
    n = 12
    g = np.linspace(-1, 1, n)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    Vs, Ts = grid.weld_mesh(*_cpu_triangle_soup(0.6 - np.sqrt(X ** 2 + Y ** 2 + Z ** 2), 0, n - 1))
    assert V.shape == Vs.shape and T.shape == Ts.shape
    de = Counter()
    for a, b, c in T:
        for e in ((a, b), (b, c), (c, a)):
            de[e] += 1
    und = Counter(tuple(sorted(e)) for e in de)
    assert max(de.values()) == 1 and set(und.values()) == {2} and len(V) - len(und) + len(T) == 2


def test_bench_clock_sampler_windows():
    """bench.py's nvidia-smi sampler: only rows stamped inside the timed region count; with fewer than two of them the
    warm-up rows (same load) are used and the line says so; throttle reasons come from the rows that were used."""
    import datetime as dt
    import importlib.util
    spec = importlib.util.spec_from_file_location("rnb_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(bench)
    finally:
        sys.argv = argv
    t = dt.datetime(2026, 1, 1, 12, 0, 0)
    ms = lambda k: dt.timedelta(milliseconds=k)
    row = lambda k, mhz, cap: [(t + ms(k)).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3], f" {mhz}", " 1965", " 900.0",
                               " Not Active", " Not Active", " Not Active", " Active" if cap else " Not Active"]
    rows = [row(0, 1965, False), row(50, 1900, False), row(100, 1700, True), row(150, 1650, True), row(200, 1600, True),
            ["garbage"], row(250, 1965, False)]
    sm, mx, reasons = bench.ClockSampler._parse(rows, t + ms(90), t + ms(210))
    assert sm == [1700.0, 1650.0, 1600.0] and mx == [1965.0] * 3 and reasons == {"sw_power_cap"}
    sm, mx, reasons = bench.ClockSampler._parse(rows, t + ms(240), t + ms(260))
    assert sm == [1965.0] and reasons == set()

    class _P:                                   # a finished sampler process
        def terminate(self): pass
        def wait(self, timeout=None): return 0
        def kill(self): pass

    def sampler(t0, t1):
        s = bench.ClockSampler.__new__(bench.ClockSampler)
        s.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        s.f.write("\n".join(",".join(r) for r in rows) + "\n")
        s.p, s.t_begin, s.t0, s.t1 = _P(), t - ms(10), t0, t1
        return s.stop()
    c = sampler(t + ms(90), t + ms(210))
    assert c["samples"] == 3 and c["sm_mhz"] == 1650.0 and c["window"] == "timed region" and c["reasons"] == ["sw_power_cap"]
    c = sampler(t + ms(110), t + ms(120))       # shorter than a sampling period: falls back to warm-up + timed rows
    assert c["samples"] == 3 and c["window"].startswith("warm-up") and c["sm_mhz"] == 1900.0
